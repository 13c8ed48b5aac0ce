"""CPU oracle (test infrastructure only -- see the module headers)."""
