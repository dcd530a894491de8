"""Test-only oracle package (see oracle/e4t_oracle.py header).  Never imported by the product path."""
