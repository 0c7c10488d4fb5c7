"""Host-side bindings of libyolo2_b200.so (the C ABI declared in include/yolo2_b200.h).

`b200.lib` loads the shared library with ctypes (and fails loudly if it is missing -- there is no
CPU fallback); `b200.ops` wraps each entry point for torch tensors (device pointers + the current
CUDA stream); `b200.engine` owns the per-shape activation plan of the Darknet-19 forward pass.
"""
