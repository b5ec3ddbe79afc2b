from oracle.oracle import line_aa  # noqa: F401  (restatement, not scikit-image)
