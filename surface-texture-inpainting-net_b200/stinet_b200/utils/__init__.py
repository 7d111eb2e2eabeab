"""Mirror of the part of the reference's `utils` package that sits next to the hot path."""
