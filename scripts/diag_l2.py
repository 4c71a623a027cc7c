import torch
p = torch.cuda.get_device_properties(0)
print(p)
import ctypes as C
rt = C.CDLL("libcudart.so.12")
class Prop(C.Structure): _fields_=[("raw", C.c_byte*4096)]
for name, attr in (("cudaDevAttrMaxPersistingL2CacheSize", 108), ("cudaDevAttrMaxAccessPolicyWindowSize", 109), ("cudaDevAttrL2CacheSize", 38)):
    v = C.c_int(); rt.cudaDeviceGetAttribute(C.byref(v), attr, 0); print(name, v.value)
for lim, name in ((5, "cudaLimitMaxL2FetchGranularity"), (6, "cudaLimitPersistingL2CacheSize")):
    v = C.c_size_t(); rt.cudaDeviceGetLimit(C.byref(v), lim); print(name, v.value)
