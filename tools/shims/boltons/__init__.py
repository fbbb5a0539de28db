"""Dev-only stand-in for `boltons` (absent offline). Just enough for gt4py.eve to import."""
