"""nncf_b200 — B200-native (sm_100a) engine for NNCF's sampling-and-scoring training loop and whole@k / given@k
evaluation, behind the reference's main.py / config surface.  Host code is Python (torch holds device memory); all
compute is in libnncf_b200.so (hand-written CUDA, C-ABI in include/nncf_b200.h).  No CPU fallback."""
__version__ = "0.1.0"
