from hf_oracle import rop as R_op  # noqa: F401
