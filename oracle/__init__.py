"""TEST INFRASTRUCTURE ONLY - the CPU checker for the CUDA product path (see oracle/fav_oracle.c)."""
