from .heterogeneous_layer_specs import (HeterogeneousLayer, get_gpt_heterogeneous_layer_spec,  # noqa: F401
                                        parse_heterogeneous_config)
