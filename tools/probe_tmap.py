"""Probe: does cuTensorMapEncodeTiled accept an overlapping-window view (stride of dim1 < extent of dim0)?"""
import torch
from cuda.bindings import driver as drv

torch.cuda.init()
x = torch.zeros(4, 28, 36, 4, device="cuda")
def enc(dims, strides, box, estr, ptr):
    r = drv.cuTensorMapEncodeTiled(drv.CUtensorMapDataType.CU_TENSOR_MAP_DATA_TYPE_FLOAT32, len(dims), ptr,
        [drv.cuuint64_t(d) for d in dims], [drv.cuuint64_t(s) for s in strides], [drv.cuuint32_t(b) for b in box],
        [drv.cuuint32_t(e) for e in estr], drv.CUtensorMapInterleave.CU_TENSOR_MAP_INTERLEAVE_NONE,
        drv.CUtensorMapSwizzle.CU_TENSOR_MAP_SWIZZLE_128B, drv.CUtensorMapL2promotion.CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
        drv.CUtensorMapFloatOOBfill.CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
    return r[0]
# window view: d0 = 32 floats, d1 = 14 windows every 32 B, d2 = 28 rows (pitch 36 px * 16 B), d3 = batch
print("overlap 4D:", enc([32, 14, 28, 4], [32, 36 * 16, 28 * 36 * 16], [32, 14, 18, 1], [1, 1, 2, 1], x.data_ptr() + 16))
print("plain   4D:", enc([4, 36, 28, 4], [16, 36 * 16, 28 * 36 * 16], [4, 8, 18, 1], [1, 1, 2, 1], x.data_ptr()))
print("unaligned base (+4B):", enc([32, 14, 28, 4], [32, 36 * 16, 28 * 36 * 16], [32, 14, 18, 1], [1, 1, 2, 1], x.data_ptr() + 4))
