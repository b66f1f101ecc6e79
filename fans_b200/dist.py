"""One process per GPU: slab bookkeeping and the NCCL communicator bootstrap for libfans_gpu (include/fans_gpu.h,
fans_comm_*).  torch.distributed is only the out-of-band channel that ships the 128-byte NCCL unique id (the role MPI_Bcast
plays for a C++ host); every data-path exchange is issued by the library itself on its own stream.

Slab sizes follow the reference (src/reader.cpp:311-331, fftw_mpi_local_size_many_transposed): rank p owns x-planes
[p*n_x/P, (p+1)*n_x/P) in real space and y rows [p*n_y/P, (p+1)*n_y/P) in Fourier space."""
import ctypes as C
import os


def slab(n, world_size, rank):
    """(start, count) of the planes rank owns along an axis of n planes."""
    if n % world_size:
        raise ValueError("axis length %d is not divisible by world_size %d" % (n, world_size))
    cnt = n // world_size
    return rank * cnt, cnt


def check_decomposition(dims, world_size):
    """The reference's constraints (reader.cpp:306, :329) + the power-of-two restriction of the CUDA FFT."""
    if world_size & (world_size - 1):
        raise ValueError("world_size must be a power of two")
    if dims[0] // 4 < world_size:
        raise ValueError("[ FANS3D_Grid ] ERROR: Number of processes too large")
    if dims[0] % world_size or dims[1] % world_size:
        raise ValueError("n_x and n_y must be divisible by world_size")


class SlabComm:
    def __init__(self, world_size, rank, handle=None):
        self.world_size, self.rank, self.handle = int(world_size), int(rank), handle

    def close(self):
        if self.handle:
            from . import _lib
            _lib.load().fans_comm_destroy(self.handle)
            self.handle = None


def broadcast_bytes(payload, src=0):
    """Ship a bytes object from rank `src` to every rank over the already initialised torch.distributed group."""
    import torch.distributed as dist
    box = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def init(device=None, backend=None):
    """Initialise torch.distributed from the torchrun environment (if needed) and create the library's NCCL communicator.
    Returns a SlabComm; with WORLD_SIZE == 1 no communicator is created."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if device is None:
        device = local
    if world == 1:
        return SlabComm(1, 0, None)
    torch.cuda.set_device(device)
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend or "nccl", device_id=torch.device("cuda", device))
    from . import _lib
    lib = _lib.load()
    buf = C.create_string_buffer(128)
    if rank == 0:
        rc = lib.fans_comm_unique_id(buf)
        if rc != 0:
            raise _lib.FansError("fans_comm_unique_id failed: " + lib.fans_last_error(None).decode())
    uid = broadcast_bytes(bytes(buf.raw) if rank == 0 else None)
    handle = C.c_void_p()
    rc = lib.fans_comm_create(C.byref(handle), world, rank, C.create_string_buffer(uid, 128), device)
    if rc != 0:
        raise _lib.FansError("fans_comm_create failed: " + lib.fans_last_error(None).decode())
    return SlabComm(world, rank, handle)
