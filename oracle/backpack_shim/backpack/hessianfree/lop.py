from hf_oracle import lop as L_op  # noqa: F401
