# The importable name of this package is ``surs_b200`` (see ../surs_b200/__init__.py).
