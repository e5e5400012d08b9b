from hf_oracle import ggnvp as ggn_vector_product_from_plist  # noqa: F401
