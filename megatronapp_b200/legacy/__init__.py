"""Pre-mcore ("legacy") model surface: ``--use-legacy-models`` keeps working, with legacy checkpoint naming."""
