from hf_oracle import hvp as hessian_vector_product  # noqa: F401
