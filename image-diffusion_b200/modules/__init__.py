"""Drop-in replacements for the reference's `modules` package (same import names, constructor signatures,
attributes and state_dict keys); all arithmetic runs in libidf_b200.so (sm_100a)."""
