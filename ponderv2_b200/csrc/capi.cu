// Library-level entry points and the dispatcher between the tensor-core and SIMT sparse-conv kernels.
#include "pv2_common.cuh"
#include <stdlib.h>
#include <string.h>

extern "C" {

int pv2_spconv_gather_gemm_simt(const void*, const void*, int64_t, int64_t, const float*, const int32_t*, const int32_t*,
                                void*, int64_t, int64_t, int, int, int, int, void*);
int pv2_spconv_wgrad_simt(const void*, const void*, const int32_t*, const int32_t*, float*, int64_t, int64_t, int, int, int,
                          int, void*);

static unsigned long long g_launches = 0;
void pv2_note_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }
int64_t pv2_launch_count(void) { return (int64_t)__atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int pv2_spconv_gather_gemm_umma(const void*, const void*, int64_t, int64_t, const float*, const int32_t*, const int32_t*,
                                void*, int64_t, int64_t, int, int, int, int, void*, size_t, void*);

// ---- runtime options (development / A-B switches).  Defaults come from the environment once; tests and benchmarks set
// them through pv2_set_option.  Names: "gg_tma" (bf16 gather through TMA gather4: -1 auto by size, 0 off, 1 on),
// "gg_bx3" (fp32 gather-GEMM as bf16x3: 0 / 1), "wgrad_mn" (MN-major bf16 weight-gradient kernel: 0 / 1).
static int g_opt_gg_tma = -2, g_opt_gg_bx3 = -2, g_opt_wgrad_mn = -2, g_opt_ksplit_max = -2, g_opt_linear_bx3 = -2;
static int g_opt_gg_groups = -2;
static int g_opt_bx3_split = -2;  // "gg_bx3_split": deep levels on the split-K bf16x3 persistent kernel (0 / 1).  Default 0: measured
// (profiles/r2v_micro_levels_*.txt) it gains 6 us at L2 64->64 and loses at L3 384->256 (74 -> 94 us) and L4 (30 -> 41 us):
// the extra weight pre-split launch and the wider reduction outweigh the cheaper chunks; the step got slower (27.9 vs 26.7 ms)   // "gg_groups": producer groups of the persistent fp32 gather-GEMM (2..4, 0 = default)
static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
int pv2_get_option(const char* name) {
  if (name == nullptr) return -1;
  if (!strcmp(name, "gg_tma")) { if (g_opt_gg_tma == -2) g_opt_gg_tma = env_int("PV2_GG_TMA", -1); return g_opt_gg_tma; }
  if (!strcmp(name, "gg_bx3")) { if (g_opt_gg_bx3 == -2) g_opt_gg_bx3 = env_int("PV2_GG_BX3", 1); return g_opt_gg_bx3; }
  if (!strcmp(name, "wgrad_mn")) { if (g_opt_wgrad_mn == -2) g_opt_wgrad_mn = env_int("PV2_WGRAD_MN", 1); return g_opt_wgrad_mn; }
  if (!strcmp(name, "linear_bx3")) { if (g_opt_linear_bx3 == -2) g_opt_linear_bx3 = env_int("PV2_LINEAR_BX3", 1); return g_opt_linear_bx3; }
  if (!strcmp(name, "gg_bx3_split")) { if (g_opt_bx3_split == -2) g_opt_bx3_split = env_int("PV2_GG_BX3_SPLIT", 0); return g_opt_bx3_split; }
  if (!strcmp(name, "gg_groups")) { if (g_opt_gg_groups == -2) g_opt_gg_groups = env_int("PV2_GG_GROUPS", 0); return g_opt_gg_groups; }
  if (!strcmp(name, "gg_ksplit_max")) { if (g_opt_ksplit_max == -2) g_opt_ksplit_max = env_int("PV2_GG_KSPLIT_MAX", 0); return g_opt_ksplit_max; }
  return -1;
}
int pv2_set_option(const char* name, int value) {
  if (name == nullptr) return PV2_EINVAL;
  if (!strcmp(name, "gg_tma")) { g_opt_gg_tma = value; return 0; }
  if (!strcmp(name, "gg_bx3")) { g_opt_gg_bx3 = value; return 0; }
  if (!strcmp(name, "wgrad_mn")) { g_opt_wgrad_mn = value; return 0; }
  if (!strcmp(name, "gg_ksplit_max")) { g_opt_ksplit_max = value; return 0; }
  if (!strcmp(name, "gg_groups")) { g_opt_gg_groups = value; return 0; }
  if (!strcmp(name, "gg_bx3_split")) { g_opt_bx3_split = value; return 0; }
  if (!strcmp(name, "linear_bx3")) { g_opt_linear_bx3 = value; return 0; }
  return PV2_EINVAL;
}

static int force_simt() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PV2_FORCE_SIMT"); v = (e && e[0] == '1') ? 1 : 0; }
  return v;
}

int pv2_version(void) { return 100; }

const char* pv2_error_string(int code) {
  if (code == 0) return "ok";
  if (code == PV2_EINVAL) return "pv2: invalid argument";
  if (code == PV2_EWORKSPACE) return "pv2: workspace too small";
  if (code == PV2_EUNSUPPORTED) return "pv2: unsupported dtype/shape";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "pv2: unknown error";
}

int pv2_sm_count(void) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return sms;
}

int pv2_spconv_gather_gemm(const void* x, const void* w, int64_t w_sco, int64_t w_sk, const float* bias,
                           const int32_t* nbr, const int32_t* row_order, void* y, int64_t n_in, int64_t n_out, int cin,
                           int cout, int kvol, int dtype, void* workspace, size_t workspace_bytes, void* stream) {
  // tensor-core implicit GEMM whenever the shape allows it (16-byte aligned channel runs, Cout <= 256, K <= 32);
  // the ragged stem (Cin = 6 / 4, K = 125) runs on the exact-fp32 SIMT kernel.  PV2_FORCE_SIMT=1 is for A/B tests.
  if (!force_simt() && (int64_t)cin * kvol >= 64) {
    int rc = pv2_spconv_gather_gemm_umma(x, w, w_sco, w_sk, bias, nbr, row_order, y, n_in, n_out, cin, cout, kvol, dtype,
                                         workspace, workspace_bytes, stream);
    if (rc != PV2_EUNSUPPORTED) return rc;
  }
  return pv2_spconv_gather_gemm_simt(x, w, w_sco, w_sk, bias, nbr, row_order, y, n_in, n_out, cin, cout, kvol, dtype, stream);
}

int pv2_wgrad_umma(const float*, int64_t, int64_t, const float*, int64_t, int64_t, const int32_t*, const int32_t*,
                   const uint8_t*, float*, int64_t, int64_t, int, int, int, void*, size_t, void*);

int pv2_wgrad_mn(const void*, const void*, const int32_t*, const int32_t*, const uint8_t*, float*, int64_t, int64_t, int, int,
                 int, int, void*);

int pv2_spconv_wgrad(const void* x, const void* dy, const int32_t* nbr, const int32_t* row_order,
                     const uint8_t* blk_active, float* dw, int64_t n_in, int64_t n_out, int cin, int cout, int kvol,
                     int dtype, void* workspace, size_t workspace_bytes, void* stream) {
  // second-generation kernel: MN-major bf16 operands (fp32 storage as bf16x3), no transposes, no scratch
  if (!force_simt() && pv2_get_option("wgrad_mn") != 0) {
    int rc = pv2_wgrad_mn(x, dy, nbr, row_order, blk_active, dw, n_in, n_out, cin, cout, kvol, dtype, stream);
    if (rc != PV2_EUNSUPPORTED) return rc;
  }
  // tensor-core wgrad pays off once the gathered rows are wide (measured on B200, 100 k voxels, K = 27:
  // 96->96 1.6 vs 2.4 ms, 256->256 6.4 vs 10.1 ms, but 32->32 1.1 vs 0.6 ms): narrow layers stay on the SIMT kernel
  static int min_cin = -1;
  if (min_cin < 0) { const char* e = getenv("PV2_WGRAD_UMMA_MIN_CIN"); min_cin = e ? atoi(e) : 8; }
  if (!force_simt() && dtype == PV2_F32 && kvol <= 128 && cin >= min_cin) {
    int rc = pv2_wgrad_umma((const float*)x, cin, 0, (const float*)dy, cout, 0, nbr, row_order, blk_active, dw, n_in, n_out,
                            cin, cout, kvol, workspace, workspace_bytes, stream);
    if (rc != PV2_EUNSUPPORTED && rc != PV2_EWORKSPACE) return rc;
  }
  return pv2_spconv_wgrad_simt(x, dy, nbr, row_order, dw, n_in, n_out, cin, cout, kvol, dtype, stream);
}

}  // extern "C"
