"""Drop-in overlay: resolves the reference import paths to the B200-native modules."""
