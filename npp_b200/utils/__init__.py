"""Mirror of the reference's `utils` package for the hot path (confusion matrix, PCKh)."""
