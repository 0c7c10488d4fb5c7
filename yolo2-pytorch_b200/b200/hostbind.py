"""Host placement for the serving path: run the launching thread -- and therefore allocate the pinned staging buffers it
touches first -- on the CPUs of the GPU's own NUMA node.

On an 8-GPU node half of the GPUs hang off each socket; a process that lands on the far socket pays a cross-socket hop on
every pinned-buffer access the copy engine makes (measured in round 1: the same end-to-end pipeline ran 17.5 k img/s on
the 8-GPU node against 25.0 k img/s on a single-GPU box).  NVML knows the CPU set next to each GPU
(`nvmlDeviceGetCpuAffinity`, what `nvidia-smi topo -m` prints); `bind_to_gpu` restricts the process to it.
"""
import os


def gpu_cpu_set(index):
    """CPUs local to GPU `index` according to NVML, or None when NVML cannot tell."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(index))
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {w * 64 + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        return cpus or None
    except Exception:
        return None


def physical_index(local_index):
    """CUDA device index -> NVML index (honours CUDA_VISIBLE_DEVICES when it lists plain integers)."""
    vis = os.environ.get('CUDA_VISIBLE_DEVICES')
    if vis:
        parts = [p.strip() for p in vis.split(',') if p.strip()]
        if local_index < len(parts) and parts[local_index].isdigit():
            return int(parts[local_index])
    return local_index


def bind_to_gpu(local_index):
    """Restrict this process to the CPUs next to its GPU.  Returns (previous affinity, new affinity or None if unchanged)."""
    try:
        before = os.sched_getaffinity(0)
    except (AttributeError, OSError):
        return None, None
    cpus = gpu_cpu_set(physical_index(local_index))
    if not cpus:
        return before, None
    target = (cpus & before) or None
    if target is None or target == before:
        return before, None
    try:
        os.sched_setaffinity(0, target)
    except OSError:
        return before, None
    return before, target


def restore(affinity):
    if affinity:
        try:
            os.sched_setaffinity(0, affinity)
        except OSError:
            pass
