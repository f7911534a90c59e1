/*
 * pl_multicast.cu -- pools on the VMM allocator and the NVLink multicast object that lets the fused kernel push a
 * finished normal tile to EVERY GPU of the box with one store (nothing in the reference: it runs on one GPU).
 *
 * The unicast push (pl_pool_attach_peers, CUDA IPC) issues one 16-byte store per peer: at 8 GPUs seven copies of every
 * tile leave through the producer's NVLink ports, which is what an all_gather costs as well.  A multicast object
 * (cuMulticastCreate) is a handle every rank binds its own physical memory to; a store to the object's mapping is
 * replicated by the NVSwitch into all bound memories -- one copy leaves the GPU.  (multimem.st and a plain st.global to
 * the multicast address are the same SASS, STG.E.128: the replication is the switch's.)
 *
 * One process per GPU: the object is created by one rank and reaches the others as a POSIX file descriptor
 * (cuMemExportToShareableHandle), which the CALLER passes between the processes (SCM_RIGHTS over a Unix socket:
 * tools/gather_tiles.py uses Python's socket.send_fds).  Driver entry points are looked up at run time
 * (cudaGetDriverEntryPoint): the library does not link libcuda and still loads on a machine without a driver.
 *
 *   every rank   pl_pool_create_shared(...)            VMM-backed pool (same kind / tile_w / capacity on all ranks)
 *   rank 0       pl_pool_mc_create(pool, n, &fd)       the multicast object for n GPUs; fd goes to the other ranks
 *   others       pl_pool_mc_import(pool, fd, n)
 *   every rank   pl_pool_mc_add_device(pool)           ... then a barrier: binding needs every device added
 *   every rank   pl_pool_mc_bind(pool)                 binds the pool's memory, maps the object; then a barrier
 *   every rank   pl_pool_push_to_peers(pool, 2)        the fused kernel's PUSH variant stores through the mapping
 */
#include <unistd.h>

#include "pl_internal.h"

namespace {

struct Drv {
    CUresult (*MemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long);
    CUresult (*MemRelease)(CUmemGenericAllocationHandle);
    CUresult (*MemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long);
    CUresult (*MemAddressFree)(CUdeviceptr, size_t);
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
    CUresult (*MemUnmap)(CUdeviceptr, size_t);
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t);
    CUresult (*MemExportToShareableHandle)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
    CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType);
    CUresult (*MulticastCreate)(CUmemGenericAllocationHandle *, const CUmulticastObjectProp *);
    CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice);
    CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long);
    CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t);
    CUresult (*MulticastGetGranularity)(size_t *, const CUmulticastObjectProp *, CUmulticastGranularity_flags);
    CUresult (*DeviceGet)(CUdevice *, int);
    CUresult (*DeviceGetAttribute)(int *, CUdevice_attribute, CUdevice);
    bool ok;
};

template <typename F> bool entry(const char *name, F *&fn)
{
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return false;
    fn = reinterpret_cast<F *>(p);
    return true;
}

const Drv &drv()
{
    static Drv d = [] {
        Drv x = {};
        x.ok = entry("cuMemCreate", x.MemCreate) && entry("cuMemRelease", x.MemRelease) && entry("cuMemAddressReserve", x.MemAddressReserve) &&
               entry("cuMemAddressFree", x.MemAddressFree) && entry("cuMemMap", x.MemMap) && entry("cuMemUnmap", x.MemUnmap) &&
               entry("cuMemSetAccess", x.MemSetAccess) && entry("cuMemExportToShareableHandle", x.MemExportToShareableHandle) &&
               entry("cuMemImportFromShareableHandle", x.MemImportFromShareableHandle) && entry("cuMulticastCreate", x.MulticastCreate) &&
               entry("cuMulticastAddDevice", x.MulticastAddDevice) && entry("cuMulticastBindMem", x.MulticastBindMem) &&
               entry("cuMulticastUnbind", x.MulticastUnbind) && entry("cuMulticastGetGranularity", x.MulticastGetGranularity) &&
               entry("cuDeviceGet", x.DeviceGet) && entry("cuDeviceGetAttribute", x.DeviceGetAttribute);
        return x;
    }();
    return d;
}

#define PL_DRV(call)                                                                                         \
    do {                                                                                                     \
        const CUresult r_ = (call);                                                                          \
        if (r_ != CUDA_SUCCESS) return pl_set_error(PL_ERR_CUDA, "%s failed (CUresult %d)", #call, (int) r_); \
    } while (0)

CUmulticastObjectProp mc_prop(int n, size_t size)
{
    CUmulticastObjectProp mp = {};
    mp.numDevices = (unsigned int) n;
    mp.size = size;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    mp.flags = 0;
    return mp;
}

CUmemAccessDesc rw_access(int device)
{
    CUmemAccessDesc a = {};
    a.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    a.location.id = device;
    a.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    return a;
}

}  // namespace

/* the pool's memory from the VMM allocator, sized to the multicast granularity (so that the whole of it can be bound) */
int pl_vmm_alloc(pl_pool *p, size_t bytes)
{
    const Drv &d = drv();
    if (!d.ok) return pl_set_error(PL_ERR_CUDA, "the driver has no virtual-memory / multicast entry points");
    CUdevice dev;
    PL_DRV(d.DeviceGet(&dev, p->ctx->device));
    int mc = 0;
    PL_DRV(d.DeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev));
    if (!mc) return pl_set_error(PL_ERR_ARG, "device %d does not support NVLink multicast", p->ctx->device);
    const CUmulticastObjectProp mp = mc_prop(2, bytes);
    size_t gran = 0;
    PL_DRV(d.MulticastGetGranularity(&gran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    const size_t size = (bytes + gran - 1) / gran * gran;
    CUmemAllocationProp ap = {};
    ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ap.location.id = p->ctx->device;
    ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    CUmemGenericAllocationHandle h = 0;
    const CUresult r = d.MemCreate(&h, size, &ap, 0);
    if (r != CUDA_SUCCESS) return pl_set_error(PL_ERR_POOL_FULL, "cuMemCreate of %zu bytes failed (CUresult %d)", size, (int) r);
    CUdeviceptr va = 0;
    if (d.MemAddressReserve(&va, size, gran, 0, 0) != CUDA_SUCCESS) {
        d.MemRelease(h);
        return pl_set_error(PL_ERR_CUDA, "cuMemAddressReserve of %zu bytes failed", size);
    }
    const CUmemAccessDesc acc = rw_access(p->ctx->device);
    if (d.MemMap(va, size, 0, h, 0) != CUDA_SUCCESS || d.MemSetAccess(va, size, &acc, 1) != CUDA_SUCCESS) {
        d.MemAddressFree(va, size);
        d.MemRelease(h);
        return pl_set_error(PL_ERR_CUDA, "mapping the pool's memory failed");
    }
    p->vmm = 1;
    p->vmm_size = size;
    p->vmm_handle = h;
    p->base = reinterpret_cast<uint8_t *>(va);
    return PL_OK;
}

void pl_vmm_free(pl_pool *p)
{
    const Drv &d = drv();
    if (!d.ok) return;
    if (p->mc_base) {
        d.MemUnmap((CUdeviceptr) p->mc_base, p->vmm_size);
        d.MemAddressFree((CUdeviceptr) p->mc_base, p->vmm_size);
    }
    if (p->mc_handle) {
        CUdevice dev;
        if (p->mc_state == 3 && d.DeviceGet(&dev, p->ctx->device) == CUDA_SUCCESS) d.MulticastUnbind(p->mc_handle, dev, 0, p->vmm_size);
        d.MemRelease(p->mc_handle);
    }
    if (p->base) {
        d.MemUnmap((CUdeviceptr) p->base, p->vmm_size);
        d.MemAddressFree((CUdeviceptr) p->base, p->vmm_size);
    }
    if (p->vmm_handle) d.MemRelease(p->vmm_handle);
    p->base = p->mc_base = nullptr;
    p->vmm_handle = p->mc_handle = 0;
    p->mc_state = 0;
}

extern "C" int pl_pool_mc_create(pl_pool *p, int n_devices, int *fd_out)
{
    if (!p || !fd_out || n_devices < 2 || n_devices > pl_pool::kMaxPeers + 1) return pl_set_error(PL_ERR_ARG, "bad argument");
    if (!p->vmm) return pl_set_error(PL_ERR_ARG, "the pool was not made by pl_pool_create_shared");
    if (p->mc_state) return pl_set_error(PL_ERR_ARG, "the pool already has a multicast object");
    const Drv &d = drv();
    PL_CUDA(cudaSetDevice(p->ctx->device));
    const CUmulticastObjectProp mp = mc_prop(n_devices, p->vmm_size);
    CUmemGenericAllocationHandle mc = 0;
    PL_DRV(d.MulticastCreate(&mc, &mp));
    int fd = -1;
    const CUresult r = d.MemExportToShareableHandle(&fd, mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
    if (r != CUDA_SUCCESS) {
        d.MemRelease(mc);
        return pl_set_error(PL_ERR_CUDA, "cuMemExportToShareableHandle failed (CUresult %d)", (int) r);
    }
    p->mc_handle = mc;
    p->mc_devices = n_devices;
    p->mc_state = 1;
    *fd_out = fd;
    return PL_OK;
}

extern "C" int pl_pool_mc_import(pl_pool *p, int fd, int n_devices)
{
    if (!p || fd < 0 || n_devices < 2 || n_devices > pl_pool::kMaxPeers + 1) return pl_set_error(PL_ERR_ARG, "bad argument");
    if (!p->vmm) return pl_set_error(PL_ERR_ARG, "the pool was not made by pl_pool_create_shared");
    if (p->mc_state) return pl_set_error(PL_ERR_ARG, "the pool already has a multicast object");
    const Drv &d = drv();
    PL_CUDA(cudaSetDevice(p->ctx->device));
    CUmemGenericAllocationHandle mc = 0;
    PL_DRV(d.MemImportFromShareableHandle(&mc, (void *) (uintptr_t) fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    p->mc_handle = mc;
    p->mc_devices = n_devices;
    p->mc_state = 1;
    return PL_OK;
}

extern "C" int pl_pool_mc_add_device(pl_pool *p)
{
    if (!p || p->mc_state != 1) return pl_set_error(PL_ERR_ARG, "no multicast object (pl_pool_mc_create / _import), or the device is added already");
    const Drv &d = drv();
    PL_CUDA(cudaSetDevice(p->ctx->device));
    CUdevice dev;
    PL_DRV(d.DeviceGet(&dev, p->ctx->device));
    PL_DRV(d.MulticastAddDevice(p->mc_handle, dev));
    p->mc_state = 2;
    return PL_OK;
}

extern "C" int pl_pool_mc_bind(pl_pool *p)
{
    if (!p || p->mc_state != 2) return pl_set_error(PL_ERR_ARG, "pl_pool_mc_add_device (and a barrier over the ranks) comes first");
    const Drv &d = drv();
    PL_CUDA(cudaSetDevice(p->ctx->device));
    PL_CUDA(cudaStreamSynchronize(p->ctx->stream));
    PL_DRV(d.MulticastBindMem(p->mc_handle, 0, p->vmm_handle, 0, p->vmm_size, 0));
    const CUmulticastObjectProp mp = mc_prop(p->mc_devices, p->vmm_size);
    size_t gran = 0;
    PL_DRV(d.MulticastGetGranularity(&gran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    CUdeviceptr va = 0;
    PL_DRV(d.MemAddressReserve(&va, p->vmm_size, gran, 0, 0));
    const CUmemAccessDesc acc = rw_access(p->ctx->device);
    if (d.MemMap(va, p->vmm_size, 0, p->mc_handle, 0) != CUDA_SUCCESS || d.MemSetAccess(va, p->vmm_size, &acc, 1) != CUDA_SUCCESS) {
        d.MemAddressFree(va, p->vmm_size);
        return pl_set_error(PL_ERR_CUDA, "mapping the multicast object failed");
    }
    p->mc_base = reinterpret_cast<uint8_t *>(va);
    p->mc_state = 3;
    return PL_OK;
}
