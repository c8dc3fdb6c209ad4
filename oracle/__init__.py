"""Oracle package — TEST INFRASTRUCTURE ONLY (see oracle/cmx_ref.py header)."""
