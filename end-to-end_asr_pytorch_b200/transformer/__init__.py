"""Drop-ins for the hot-path members of the reference package `transformer`
(/root/reference/src/transformer): attention.py, cif_model.py, loss.py."""
