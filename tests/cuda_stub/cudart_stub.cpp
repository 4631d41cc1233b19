// Recording CUDA-runtime stub (TEST INFRASTRUCTURE, never shipped): the product's host objects (plan.o, engine.o)
// linked against this file instead of libcudart run their whole launch logic on a machine without a GPU.  "Device"
// memory is host memory, kernels are not executed; every launch (name, grid, block, shared memory, stream, cluster
// size, the UmmaArgs of the tcgen05 kernels and the tensor maps they were given), every tensor-map encode and every
// stream / event operation is recorded for tests/test_host_launch.py, which checks the launch invariants the kernels
// rely on (shared-memory and TMEM budgets, TMA descriptor rules, tensor extents inside the workspace, fork / join
// structure of the side streams) for all architectures, batch sizes and library switches.
#include <cuda.h>
#include <cuda_runtime_api.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../vae_npvc_b200/csrc/launch_args.h"

extern "C" {

struct StubTmap {
  int32_t dtype, rank;
  uint64_t base;
  uint64_t dims[5], strides[5];
  uint32_t box[5], estrides[5];
  int32_t interleave, swizzle, l2promo, oob;
};
struct StubUmma {
  int32_t K, N, BN, kblocks, sw, stages, tmem_cols;
  int32_t Rb, Ra, Ab, FB, TA, RbH, rows_tile, frames, m_tiles;
  int32_t n_tiles, acc_sets, tapT, tapC, tapP, b_tile_al, d_sw, rows_al, tiles_per_split, ld;
  uint64_t c_ptr; int64_t c_fs; int32_t c_R, c_rs, c_off, c_flen, c_pred, c_split;
  uint64_t out_ptr;
  int32_t a_boxes, ln_on, ln_store_c, ln_L, ln_Cn, ln_out_flen, ln_out_off, b_res; uint64_t ln_aout;
};
struct StubLaunch {
  char name[192];
  uint32_t grid[3], block[3];
  uint64_t smem, stream, seq;
  uint32_t cluster_x; int32_t has_umma;
  int32_t tmap[4];
  StubUmma u;
};
struct StubOp { int32_t kind; uint64_t stream, event, seq, ptr, bytes; };   // 0 launch, 1 event record, 2 stream wait, 3 memset, 4 memcpy

}  // extern "C"

namespace {
std::map<const void*, std::string> g_names;
std::vector<StubLaunch> g_launches;
std::vector<StubTmap> g_tmaps;
std::vector<StubOp> g_ops;
uint64_t g_seq = 0, g_next_handle = 0x1000;
int g_fail_cluster = 0, g_devices = 1;
cudaError_t g_last = cudaSuccess;
struct CallCfg { dim3 grid, block; size_t smem; void* stream; };
thread_local std::vector<CallCfg> g_cfg;
const uint64_t TMAP_MAGIC = 0x5354554250414d54ull;      // "TMAPBUTS"

void log_op(int kind, uint64_t stream, uint64_t event, uint64_t ptr = 0, uint64_t bytes = 0) {
  g_ops.push_back(StubOp{kind, stream, event, g_seq++, ptr, bytes});
}

CUresult stub_encode(CUtensorMap* tm, CUtensorMapDataType dt, cuuint32_t rank, void* addr, const cuuint64_t* gdim, const cuuint64_t* gstride,
                     const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapInterleave il, CUtensorMapSwizzle sw, CUtensorMapL2promotion l2,
                     CUtensorMapFloatOOBfill oob) {
  if (!tm || rank < 1 || rank > 5) return CUDA_ERROR_INVALID_VALUE;
  StubTmap t; memset(&t, 0, sizeof t);
  t.dtype = (int)dt; t.rank = (int)rank; t.base = (uint64_t)(uintptr_t)addr;
  for (cuuint32_t i = 0; i < rank; i++) { t.dims[i] = gdim[i]; t.box[i] = box[i]; t.estrides[i] = estr[i]; if (i + 1 < rank) t.strides[i] = gstride[i]; }
  t.interleave = (int)il; t.swizzle = (int)sw; t.l2promo = (int)l2; t.oob = (int)oob;
  memset(tm, 0, sizeof(CUtensorMap));
  uint64_t tag[2] = {TMAP_MAGIC, (uint64_t)g_tmaps.size()};
  memcpy(tm, tag, sizeof tag);
  g_tmaps.push_back(t);
  return CUDA_SUCCESS;
}

void record_launch(const void* func, dim3 grid, dim3 block, void** args, size_t smem, cudaStream_t stream, unsigned cluster_x) {
  StubLaunch L; memset(&L, 0, sizeof L);
  auto it = g_names.find(func);
  snprintf(L.name, sizeof L.name, "%s", it == g_names.end() ? "?" : it->second.c_str());
  L.grid[0] = grid.x; L.grid[1] = grid.y; L.grid[2] = grid.z; L.block[0] = block.x; L.block[1] = block.y; L.block[2] = block.z;
  L.smem = smem; L.stream = (uint64_t)(uintptr_t)stream; L.seq = g_seq; L.cluster_x = cluster_x;
  for (int i = 0; i < 4; i++) L.tmap[i] = -1;
  if (strstr(L.name, "umma_fwd_kernel_t") || strstr(L.name, "umma_wgrad_kernel_t")) {
    L.has_umma = 1;
    for (int i = 0; i < 4; i++) {
      uint64_t tag[2]; memcpy(tag, args[i], sizeof tag);
      if (tag[0] == TMAP_MAGIC) L.tmap[i] = (int32_t)tag[1];
    }
    const npvc::UmmaArgs& g = *reinterpret_cast<const npvc::UmmaArgs*>(args[4]);
    StubUmma& u = L.u;
    u.K = g.K; u.N = g.N; u.BN = g.BN; u.kblocks = g.kblocks; u.sw = g.sw; u.stages = g.stages; u.tmem_cols = g.tmem_cols;
    u.Rb = g.rt.Rb; u.Ra = g.rt.Ra; u.Ab = g.rt.Ab; u.FB = g.rt.FB; u.TA = g.rt.TA; u.RbH = g.rt.RbH; u.rows_tile = g.rt.rows_tile;
    u.frames = g.rt.frames; u.m_tiles = g.rt.m_tiles; u.n_tiles = g.n_tiles; u.acc_sets = g.acc_sets; u.tapT = g.tapT; u.tapC = g.tapC;
    u.tapP = g.tapP; u.b_tile_al = g.b_tile_al; u.d_sw = g.d_sw; u.rows_al = g.rows_al; u.tiles_per_split = g.tiles_per_split; u.ld = g.ld;
    u.a_boxes = g.a_boxes; u.b_res = g.b_res; u.ln_on = g.ln.on; u.ln_store_c = g.ln.store_c; u.ln_L = g.ln.L; u.ln_Cn = g.ln.Cn; u.ln_out_flen = g.ln.out_flen; u.ln_out_off = g.ln.out_off;
    u.ln_aout = (uint64_t)(uintptr_t)g.ln.aout;
    u.c_ptr = (uint64_t)(uintptr_t)g.C.p; u.c_fs = g.C.fs; u.c_R = g.C.R; u.c_rs = g.C.rs; u.c_off = g.C.off; u.c_flen = g.C.flen;
    u.c_pred = g.C.pred; u.c_split = g.C.split; u.out_ptr = (uint64_t)(uintptr_t)g.out;
  }
  g_launches.push_back(L);
  log_op(0, L.stream, 0);
}
}  // namespace

extern "C" {

// ---- test-side API ---------------------------------------------------------------------------
void stub_reset(void) { g_launches.clear(); g_tmaps.clear(); g_ops.clear(); g_last = cudaSuccess; }
int64_t stub_n_launches(void) { return (int64_t)g_launches.size(); }
int stub_get_launch(int64_t i, StubLaunch* out) { if (i < 0 || i >= (int64_t)g_launches.size()) return 1; *out = g_launches[i]; return 0; }
int64_t stub_n_tmaps(void) { return (int64_t)g_tmaps.size(); }
int stub_get_tmap(int64_t i, StubTmap* out) { if (i < 0 || i >= (int64_t)g_tmaps.size()) return 1; *out = g_tmaps[i]; return 0; }
int64_t stub_n_ops(void) { return (int64_t)g_ops.size(); }
int stub_get_op(int64_t i, StubOp* out) { if (i < 0 || i >= (int64_t)g_ops.size()) return 1; *out = g_ops[i]; return 0; }
void stub_fail_cluster_launches(int n) { g_fail_cluster = n; }      // the next n cluster launches are refused
void stub_set_device_count(int n) { g_devices = n; }

// ---- registration (called by the static constructors nvcc generates in engine.o) -----------------
void** __cudaRegisterFatBinary(void*) { static void* handle = nullptr; return &handle; }
void __cudaRegisterFatBinaryEnd(void**) {}
void __cudaUnregisterFatBinary(void**) {}
void __cudaRegisterFunction(void**, const char* hostFun, char*, const char* deviceName, int, uint3*, uint3*, dim3*, dim3*, int*) {
  g_names[(const void*)hostFun] = deviceName;
}
unsigned __cudaPushCallConfiguration(dim3 grid, dim3 block, size_t smem, struct CUstream_st* stream) {
  g_cfg.push_back(CallCfg{grid, block, smem, (void*)stream}); return 0;
}
cudaError_t __cudaPopCallConfiguration(dim3* grid, dim3* block, size_t* smem, void* stream) {
  if (g_cfg.empty()) return cudaErrorInvalidConfiguration;
  CallCfg c = g_cfg.back(); g_cfg.pop_back();
  *grid = c.grid; *block = c.block; *smem = c.smem; *(void**)stream = c.stream; return cudaSuccess;
}

// ---- runtime ------------------------------------------------------------------------------------
cudaError_t cudaGetDeviceCount(int* n) { *n = g_devices; return g_devices ? cudaSuccess : cudaErrorNoDevice; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int* v, enum cudaDeviceAttr a, int) { *v = (a == cudaDevAttrMultiProcessorCount) ? 148 : 0; return cudaSuccess; }
cudaError_t cudaGetDriverEntryPoint(const char* sym, void** fn, unsigned long long, enum cudaDriverEntryPointQueryResult* st) {
  if (!strcmp(sym, "cuTensorMapEncodeTiled")) { *fn = (void*)&stub_encode; if (st) *st = cudaDriverEntryPointSuccess; return cudaSuccess; }
  *fn = nullptr; if (st) *st = cudaDriverEntryPointSymbolNotFound; return cudaSuccess;
}
const char* cudaGetErrorString(cudaError_t e) { static char b[32]; snprintf(b, sizeof b, "stub error %d", (int)e); return b; }
cudaError_t cudaGetLastError(void) { cudaError_t e = g_last; g_last = cudaSuccess; return e; }
cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, enum cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, enum cudaMemcpyKind, cudaStream_t st) {
  log_op(4, (uint64_t)(uintptr_t)st, 0, (uint64_t)(uintptr_t)d, n); memmove(d, s, n); return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t hgt, enum cudaMemcpyKind, cudaStream_t st) {
  log_op(4, (uint64_t)(uintptr_t)st, 0, (uint64_t)(uintptr_t)d, hgt ? (hgt - 1) * dp + w : 0);
  for (size_t r = 0; r < hgt; r++) memmove((char*)d + r * dp, (const char*)s + r * sp, w);
  return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t st) {
  log_op(3, (uint64_t)(uintptr_t)st, 0, (uint64_t)(uintptr_t)d, n); memset(d, v, n); return cudaSuccess;
}
cudaError_t cudaMemset2DAsync(void* d, size_t pitch, int v, size_t w, size_t hgt, cudaStream_t st) {
  log_op(3, (uint64_t)(uintptr_t)st, 0, (uint64_t)(uintptr_t)d, hgt ? (hgt - 1) * pitch + w : 0);      // extent touched: the workspace-bounds checks see it
  for (size_t r = 0; r < hgt; r++) memset((char*)d + r * pitch, v, w);
  return cudaSuccess;
}
cudaError_t cudaFuncSetAttribute(const void*, enum cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)(uintptr_t)(g_next_handle++); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)(uintptr_t)(g_next_handle++); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) { log_op(1, (uint64_t)(uintptr_t)s, (uint64_t)(uintptr_t)e); return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned) { log_op(2, (uint64_t)(uintptr_t)s, (uint64_t)(uintptr_t)e); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.001f; return cudaSuccess; }
cudaError_t cudaLaunchKernel(const void* f, dim3 grid, dim3 block, void** args, size_t smem, cudaStream_t st) {
  record_launch(f, grid, block, args, smem, st, 1); return cudaSuccess;
}
cudaError_t cudaLaunchKernelExC(const cudaLaunchConfig_t* c, const void* f, void** args) {
  unsigned cx = 1;
  for (unsigned i = 0; i < c->numAttrs; i++) if (c->attrs[i].id == cudaLaunchAttributeClusterDimension) cx = c->attrs[i].val.clusterDim.x;
  if (cx > 1 && g_fail_cluster > 0) { g_fail_cluster--; g_last = cudaErrorLaunchOutOfResources; return cudaErrorLaunchOutOfResources; }
  record_launch(f, c->gridDim, c->blockDim, args, c->dynamicSmemBytes, c->stream, cx); return cudaSuccess;
}

}  // extern "C"
