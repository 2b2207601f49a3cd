"""Reference arm helpers (measurement / test infrastructure, never on the product path)."""
