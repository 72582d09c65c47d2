"""Test-only CPU oracle (see tt_oracle.py header).  Never imported by the product path."""
