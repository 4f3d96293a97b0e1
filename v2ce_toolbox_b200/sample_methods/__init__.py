"""Stage-2 comparison samplers (the reference keeps them under train/scripts/stage2/sample_methods/)."""
