"""B200-native GP latent force model hot path (drop-in for wejpurvis/DIS_project's src/ modules)."""
__version__ = "0.1.0"
