"""unused by the code path (imported only)."""
