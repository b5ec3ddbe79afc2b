"""Stand-in package: routes skimage.draw.line_aa / skimage.morphology.remove_small_holes to the
oracle's RESTATEMENTS (oracle/mpc_oracle.c::orc_line_aa, oracle/oracle.py::remove_small_holes).
This is NOT scikit-image; anything produced through it is 'restated-dependency' data."""
__version__ = "0+oracle-restatement"
