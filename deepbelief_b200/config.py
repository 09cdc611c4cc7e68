"""Library-wide dtypes (reference: deepbelief/config.py:1-2 — everything is fp32 / int32)."""
float_type = 'float32'
int_type = 'int32'
