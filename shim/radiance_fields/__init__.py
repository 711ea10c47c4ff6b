"""`radiance_fields` of the reference (its __init__.py is empty) -> the B200 product's modules."""
