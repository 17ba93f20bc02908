// Host-side launch helpers shared by the kernel translation units: everything that is cached is cached PER DEVICE
// (a process may hold handles on several GPUs), and TMA descriptors are encoded once per (buffer, geometry).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

// multiprocessor count of the CURRENT device
int tc_num_sms();
// cudaFuncSetAttribute(func, MaxDynamicSharedMemorySize, bytes), issued once per (current device, func)
cudaError_t tc_func_smem(const void* func, int bytes);
// 5-D tiled map over packed [2B][E][E][E][64] fp16 planes (hi planes then lo planes), box {64, bz, by, bx, 1},
// SWIZZLE_128B; encoded once per (device, base, B, E, by, bz) and copied from the cache afterwards
bool tc_make_act_map(CUtensorMap* map, const __half* base, int B, int E, int by, int bz, int bx = 1);
// drop every cached descriptor that points into [base, base + bytes) (call before freeing device memory that was
// used as a TMA source, so a recycled address never meets a stale descriptor with another geometry)
void tc_forget_maps(const void* base, size_t bytes);
// 3-D tiled map over the two packed planes of an Act viewed as a matrix of voxel rows: {64 channels, nrows, 2 planes},
// box {64, box_rows, 1}, SWIZZLE_128B (plane 0 = hi, plane 1 = lo; `plane_elems` = distance between the planes in halves)
bool tc_make_rows_map(CUtensorMap* map, const __half* base, long long nrows, long long plane_elems, int box_rows);
// 5-D tiled fp32 map over a planar gradient [3][B][D][D][D] ({z, y, x, b, c}), box {D + 8, by, 3, 1, 1} (z from -4: box starts must be 16-byte aligned), no swizzle: the
// head backward stages a tile's neighbourhood with it (coordinates may start below 0; out-of-bounds elements arrive as 0)
bool tc_make_gplanar_map(CUtensorMap* map, const float* base, int B, int D, int by);
