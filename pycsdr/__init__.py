"""Drop-in `pycsdr` package backed by libowrx_b200.so."""
