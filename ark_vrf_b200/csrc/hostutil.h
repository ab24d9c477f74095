// Host-side helpers of libavrf_gpu.so (hostutil.cpp).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace avrf {
// Copy into the pinned staging buffers of single pushes with non-temporal stores (32-byte multiples into 32-byte
// aligned destinations; plain memcpy otherwise): the staged bytes are read next by the GPU's DMA engine, not by
// this core, so they need not displace the caller's working set or cost a read-for-ownership.
void stage_copy(void* dst, const void* src, size_t n);
// One proof of a single push into the staging arrays (all destinations 32-byte aligned), one call.
void stage_proof(uint8_t* dpk, uint8_t* dr, uint8_t* ds, uint8_t* dio, const uint8_t* pk, const uint8_t* r, const uint8_t* s,
                 const uint8_t* ios, size_t io_bytes);
// Orders the non-temporal stores before the DMA is queued.
void stage_fence();
}  // namespace avrf
