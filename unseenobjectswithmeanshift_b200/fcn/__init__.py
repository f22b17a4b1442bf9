"""Mirror of the two-stage test-time glue of the reference's lib/fcn package (test_dataset.py)."""
