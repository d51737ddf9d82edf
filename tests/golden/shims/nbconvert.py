"""Empty stand-in (imported by the reference analysis package, never used on the golden path)."""
