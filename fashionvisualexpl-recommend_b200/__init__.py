"""B200-native BPR train step + full-catalog top-k evaluation (BPRMF / VBPR).

Drop-in for the hot path of peternara/FashionVisualExpl-recommend: the host side
mirrors the reference's Python protocol (``DataLoader``, ``BPRMF``, ``VBPR``,
``Evaluator``, ``train_rec``), and every numeric operation goes through the C-ABI
shared library ``csrc/libfvx.so`` (hand-written sm_100a CUDA, declared in
``include/fvx.h``).  There is no CPU fallback: device calls raise if the library
or a CUDA device is missing.
"""
__version__ = "0.1.0"
