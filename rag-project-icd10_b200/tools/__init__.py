"""Drop-in mirror of the reference's tools/ on the hot path (build_database.py)."""
