from oracle.oracle import remove_small_holes  # noqa: F401  (restatement, not scikit-image)
