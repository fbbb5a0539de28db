"""Dev-only stand-in for `mako` (absent offline); only the GridTools C++ codegen renders Mako templates."""
