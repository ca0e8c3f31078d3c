"""Mirror of the reference's `models` package (operations, genotypes, model_augment,
model_search_interact) on libnpp_b200 kernels."""
