"""Drop-in replacement of the reference's `model` package (model/dgcnn.py, model/attention.py, model/capl.py).

Put `gfs-3dseg_gws_b200/` on PYTHONPATH ahead of the reference checkout and `train.py`, `runs/eval.py` and
`get_basis.py` import these modules unchanged.  Constructors, forward signatures, return conventions and state-dict
keys are the reference's; the arithmetic runs in the hand-written sm_100a kernels behind include/gfs3d.h.
"""
