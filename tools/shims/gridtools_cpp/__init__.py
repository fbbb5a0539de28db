"""Dev-only stand-in: GridTools headers are not available offline."""


def get_include_dir():
    return "/nonexistent/gridtools/include"
