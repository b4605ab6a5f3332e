"""Host-side helpers mirrored from the reference package `utils`."""
