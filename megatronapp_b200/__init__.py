"""megatronapp_b200: a B200-native (sm_100a) 3D-parallel LLM training framework with the
capabilities of OpenSQZ/MegatronApp (MegaScan / MegaFBD / MegaDPP / MegaScope)."""
__version__ = "0.1.0"
