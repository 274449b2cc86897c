"""CPU oracle (test infrastructure only) - see ctclip_oracle.py."""
