"""Feasibility probe for the multicast push: N processes, one per GPU of one box, share ONE NVLink multicast
object through a POSIX file descriptor passed over a pipe (SCM_RIGHTS); every process binds its own VMM allocation to it;
process 0 stores through the multicast mapping (mc_store.cubin: multimem.st = STG to the multicast address) and every
process finds the words in its own memory.
    nvcc -gencode arch=compute_100a,code=sm_100a -cubin -o tools/microbench/mc_store.cubin tools/microbench/mc_store.cu
    python tools/microbench/mc_push_probe.py [N]"""
import multiprocessing as mp
import multiprocessing.reduction as red
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def chk(r):
    from cuda import cuda
    err = r[0]
    if err != cuda.CUresult.CUDA_SUCCESS:
        raise RuntimeError(str(err))
    return r[1] if len(r) == 2 else (r[1:] if len(r) > 2 else None)


def worker(rank, n, conns, barrier, out):
    from cuda import cuda
    try:
        chk(cuda.cuInit(0))
        dev = chk(cuda.cuDeviceGet(rank))
        ctx = chk(cuda.cuDevicePrimaryCtxRetain(dev))
        chk(cuda.cuCtxSetCurrent(ctx))
        FD = cuda.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
        mcp = cuda.CUmulticastObjectProp()
        mcp.numDevices = n
        mcp.handleTypes = FD
        mcp.flags = 0
        want = 64 << 20
        mcp.size = want
        gran = chk(cuda.cuMulticastGetGranularity(mcp, cuda.CUmulticastGranularity_flags.CU_MULTICAST_GRANULARITY_RECOMMENDED))
        size = (want + gran - 1) // gran * gran
        mcp.size = size
        if rank == 0:
            mc = chk(cuda.cuMulticastCreate(mcp))
            fd = chk(cuda.cuMemExportToShareableHandle(mc, FD, 0))
            for r in range(1, n):
                red.send_handle(conns[r], int(fd), None)
        else:
            fd = red.recv_handle(conns[rank])
            mc = chk(cuda.cuMemImportFromShareableHandle(fd, FD))
        chk(cuda.cuMulticastAddDevice(mc, dev))
        barrier.wait()
        prop = cuda.CUmemAllocationProp()
        prop.type = cuda.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
        prop.location.type = cuda.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        prop.location.id = rank
        prop.requestedHandleTypes = FD
        mem = chk(cuda.cuMemCreate(size, prop, 0))
        chk(cuda.cuMulticastBindMem(mc, 0, mem, 0, size, 0))
        acc = cuda.CUmemAccessDesc()
        acc.location.type = cuda.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        acc.location.id = rank
        acc.flags = cuda.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
        va = chk(cuda.cuMemAddressReserve(size, gran, 0, 0))
        chk(cuda.cuMemMap(va, size, 0, mem, 0))
        chk(cuda.cuMemSetAccess(va, size, [acc], 1))
        mva = chk(cuda.cuMemAddressReserve(size, gran, 0, 0))
        chk(cuda.cuMemMap(mva, size, 0, mc, 0))
        chk(cuda.cuMemSetAccess(mva, size, [acc], 1))
        chk(cuda.cuMemsetD8(va, 0, size))
        chk(cuda.cuCtxSynchronize())
        barrier.wait()
        nwords = 1 << 20
        if rank == 0:
            mod = chk(cuda.cuModuleLoadData(open(os.path.join(HERE, "mc_store.cubin"), "rb").read()))
            fn = chk(cuda.cuModuleGetFunction(mod, b"mc_store"))
            args = (np.array([int(mva)], np.uint64), np.array([nwords], np.int32), np.array([7.0], np.float32))
            ptrs = np.array([a.ctypes.data for a in args], np.uint64)
            e0, e1 = chk(cuda.cuEventCreate(0)), chk(cuda.cuEventCreate(0))
            chk(cuda.cuLaunchKernel(fn, nwords // 256, 1, 1, 256, 1, 1, 0, 0, ptrs.ctypes.data, 0))
            chk(cuda.cuCtxSynchronize())
            chk(cuda.cuEventRecord(e0, 0))
            for _ in range(10):
                chk(cuda.cuLaunchKernel(fn, nwords // 256, 1, 1, 256, 1, 1, 0, 0, ptrs.ctypes.data, 0))
            chk(cuda.cuEventRecord(e1, 0))
            chk(cuda.cuCtxSynchronize())
            ms = chk(cuda.cuEventElapsedTime(e0, e1)) / 10
            out.put(("time", rank, "%.3f ms per 16 MiB multicast store = %.0f GB/s delivered to each of %d GPUs" % (ms, 16.777216 / ms, n)))
        barrier.wait()
        host = np.zeros(nwords * 4, np.float32)
        chk(cuda.cuMemcpyDtoH(host.ctypes.data, va, host.nbytes))
        h = host.reshape(-1, 4)
        ok = bool((h[:, 0] == 7.0).all() and (h[:, 1] == np.arange(nwords, dtype=np.float32)).all() and (h[:, 2] == 8.0).all())
        out.put(("result", rank, ok))
        barrier.wait()
    except Exception as ex:
        out.put(("error", rank, repr(ex)))
        try:
            barrier.abort()
        except Exception:
            pass


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    mp.set_start_method("spawn")
    barrier = mp.Barrier(n)
    out = mp.Queue()
    pipes = [mp.Pipe() for _ in range(n)]
    procs = []
    for r in range(n):
        conns = [p[0] for p in pipes] if r == 0 else [p[1] for p in pipes]
        procs.append(mp.Process(target=worker, args=(r, n, conns, barrier, out)))
        procs[-1].start()
    for p in procs:
        p.join(120)
    while not out.empty():
        print(out.get())
