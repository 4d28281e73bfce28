"""oracle — TEST INFRASTRUCTURE, NOT PRODUCT (see oracle/qsim_oracle.c header)."""
