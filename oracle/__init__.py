"""Test infrastructure: CPU restatement of the reference's hot path.  Never imported by the product."""
