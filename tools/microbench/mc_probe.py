from cuda import cuda
def chk(r):
    if isinstance(r, tuple):
        err = r[0]
        if err != cuda.CUresult.CUDA_SUCCESS: raise RuntimeError(str(err))
        return r[1] if len(r) == 2 else r[1:]
chk(cuda.cuInit(0))
n = chk(cuda.cuDeviceGetCount())
print("devices", n)
A = cuda.CUdevice_attribute
for name in ("CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED", "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED",
             "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED", "CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED",
             "CU_DEVICE_ATTRIBUTE_GPU_DIRECT_RDMA_WITH_CUDA_VMM_SUPPORTED"):
    try:
        print(name, chk(cuda.cuDeviceGetAttribute(getattr(A, name), 0)))
    except Exception as e:
        print(name, "error", e)
