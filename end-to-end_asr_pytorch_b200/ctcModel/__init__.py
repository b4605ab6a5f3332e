"""Drop-ins for the hot-path members of the reference package `ctcModel`
(/root/reference/src/ctcModel): attention.py, loss.py."""
