"""Mirror of the reference's `core` package for the hot path (criterion, evaluate)."""
