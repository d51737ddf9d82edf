"""Empty stand-in for matplotlib (only imported, never used, on the golden-generation path)."""
