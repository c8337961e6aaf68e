// tc_layout.cuh -- bf16 tile-contiguous X caches, factor packing, TMA tensor maps, the Factor record, launch helper
// Part of the tensor-core engine; included only by tc_engine.cu, inside namespace nmfb200 and its
// anonymous namespace.
#pragma once

// ---- X caches: bf16, TILE-CONTIGUOUS ---------------------------------------------------------------------
// A panel with R rows and contraction length Kdim is stored as [tile][kb][TR rows][64 cols] (TR = rows per
// CTA, kb = 64-wide k-block): one TMA box = one contiguous TR*128-byte burst and a CTA streams one
// sequential region of HBM (instead of gathering 128-B segments from TR rows a full row pitch apart).
// Padding rows / columns are written as zeros.  element(r, c) = X[r*sr + c*sc].
// (a) contraction index contiguous in the source (sc == 1): direct
// dst_lo (optional): the bf16 remainder X - bf16(X) in the same layout (precision mode bf16x3)
__global__ void cvt_tiled_direct_kernel(const float* __restrict__ X, int64_t sr, int R, int Kdim, int TR, int nkb,
                                        bf16* __restrict__ dst, bf16* __restrict__ dst_lo) {
    const int64_t row_slot = blockIdx.x;            // tile * TR + row-in-tile (x: up to 2^31-1 rows)
    const int tile = (int)(row_slot / TR), rr = (int)(row_slot % TR);
    const int64_t r = (int64_t)tile * TR + rr;
    for (int c = blockIdx.y * blockDim.x + threadIdx.x; c < nkb * 64; c += gridDim.y * blockDim.x) {
        float v = (r < R && c < Kdim) ? X[r * sr + c] : 0.f;
        const int kb = c >> 6, cc = c & 63;
        const bf16 hi = __float2bfloat16_rn(v);
        const int64_t o = (((int64_t)tile * nkb + kb) * TR + rr) * 64 + cc;
        dst[o] = hi;
        if (dst_lo != nullptr) dst_lo[o] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
}
// (b) row index contiguous in the source (sr == 1): 64 x 64 transpose through shared memory
__global__ void cvt_tiled_transpose_kernel(const float* __restrict__ X, int64_t sc, int R, int Kdim, int TR, int nkb, int tiles,
                                           bf16* __restrict__ dst, bf16* __restrict__ dst_lo) {
    __shared__ float tile_s[64][65];
    const int kb = blockIdx.x;
    const int64_t r_base = (int64_t)blockIdx.y * 64;   // 64 consecutive logical rows
    for (int y = threadIdx.y; y < 64; y += blockDim.y) {  // y: column within the k-block, x: row (contiguous in X)
        int64_t r = r_base + threadIdx.x * 2;
        int c = kb * 64 + y;
        float v0 = (r < R && c < Kdim) ? X[r + (int64_t)c * sc] : 0.f;
        float v1 = (r + 1 < R && c < Kdim) ? X[r + 1 + (int64_t)c * sc] : 0.f;
        tile_s[threadIdx.x * 2][y] = v0;
        tile_s[threadIdx.x * 2 + 1][y] = v1;
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 64; y += blockDim.y) {  // y: row within the 64-row group, x: column pair
        int64_t r = r_base + y;
        const int tile = (int)(r / TR), rr = (int)(r % TR);
        if (tile >= tiles) continue;  // the grid is rounded up to 64-row groups
        const float v0 = tile_s[y][threadIdx.x * 2], v1 = tile_s[y][threadIdx.x * 2 + 1];
        __nv_bfloat162 pk = __floats2bfloat162_rn(v0, v1);
        const int64_t o = (((int64_t)tile * nkb + kb) * TR + rr) * 64 + threadIdx.x * 2;
        *(__nv_bfloat162*)(dst + o) = pk;
        if (dst_lo != nullptr)
            *(__nv_bfloat162*)(dst_lo + o) = __floats2bfloat162_rn(v0 - __bfloat162float(pk.x), v1 - __bfloat162float(pk.y));
    }
}

// ---- factor packing / unpacking -----------------------------------------------------------------------
// src(r, a) = S[r*sr + a*sa] (r < R, a < k) -> Fm[r][a], Fhi, Flo ([R][KP]) and FbT[a][r] ([KP][ldT]); zero padded
__global__ void pack_factor_kernel(const float* __restrict__ S, int64_t sr, int64_t sa, int R, int k, int KP, float* __restrict__ Fm,
                                   bf16* __restrict__ Fhi, bf16* __restrict__ Flo, bf16* __restrict__ FbT, int64_t ldT) {
    const int64_t total = (int64_t)R * KP;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = idx / KP;
        const int a = (int)(idx % KP);
        float v = a < k ? S[r * sr + a * sa] : 0.f;
        bf16 hi = __float2bfloat16_rn(v);
        Fm[idx] = v;
        Fhi[idx] = hi;
        Flo[idx] = __float2bfloat16_rn(v - __bfloat162float(hi));
        FbT[(int64_t)a * ldT + r] = hi;
    }
}
// FbTlo[a][r] = Flo[r][a]: the transposed copy of the bf16 remainder (B operand of the hi*lo pass, precision mode bf16x3)
__global__ void transpose_lo_kernel(const bf16* __restrict__ Flo, int R, int KP, bf16* __restrict__ FbTlo, int64_t ldT) {
    __shared__ unsigned short t[32][33];
    const unsigned short* src = (const unsigned short*)Flo;
    unsigned short* dst = (unsigned short*)FbTlo;
    const int r0 = blockIdx.x * 32, a0 = blockIdx.y * 32;
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int r = r0 + y, a = a0 + threadIdx.x;
        t[y][threadIdx.x] = (r < R) ? src[(size_t)r * KP + a] : (unsigned short)0;
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int a = a0 + y, r = r0 + threadIdx.x;
        if (r < R) dst[(size_t)a * ldT + r] = t[threadIdx.x][y];
    }
}
__global__ void unpack_factor_kernel(const float* __restrict__ Fm, int R, int k, int KP, float* __restrict__ D, int64_t sr, int64_t sa) {
    const int64_t total = (int64_t)R * k;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = idx / k;
        const int a = (int)(idx % k);
        D[r * sr + a * sa] = Fm[r * KP + a];
    }
}

// ---- host side ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        NMF_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        NMF_REQUIRE(p != nullptr && qres == cudaDriverEntryPointSuccess, NMFB200_ECUDA, "cuTensorMapEncodeTiled unavailable");
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

// bf16 matrix [rows][inner] with row pitch ld (elements); box = 64 (128 B) x box_rows, SWIZZLE_128B
CUtensorMap make_tmap_bf16(const void* ptr, uint64_t inner, uint64_t rows, uint64_t ld, uint32_t box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {ld * sizeof(bf16)};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    NMF_REQUIRE(r == CUDA_SUCCESS, NMFB200_ECUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return m;
}

// fp32 matrix [rows][inner] with row pitch ld (elements); box = 32 (128 B) x box_rows, SWIZZLE_128B
CUtensorMap make_tmap_f32(const void* ptr, uint64_t inner, uint64_t rows, uint64_t ld, uint32_t box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {ld * sizeof(float)};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    NMF_REQUIRE(r == CUDA_SUCCESS, NMFB200_ECUDA, "cuTensorMapEncodeTiled(f32) failed with code " + std::to_string((int)r));
    return m;
}

inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
inline int pick_kp(int64_t k) { return k <= 64 ? 64 : (k <= 128 ? 128 : (k <= 256 ? 256 : 0)); }
inline int ew_grid(int64_t len) { return (int)std::min<int64_t>(ceil_div(len, 256), 148 * 16); }

struct Factor {  // one factor in row-factor layout
    int R = 0;
    int rowsT = 0;
    int64_t ldT = 0;
    float* m = nullptr;
    bf16 *hi = nullptr, *lo = nullptr, *bT = nullptr;
    bf16* bTlo = nullptr;  // precision mode bf16x3 only: transposed copy of lo
    float* P = nullptr;  // Gram of THIS factor (k x k), fp32 accumulator
    bf16 *Phi = nullptr, *Plo = nullptr;
    float* conv = nullptr;
    float* colsum = nullptr;  // [KP] column sums (MultUpdate :div)
    int tiles = 0;
    int tile_rows = 128;
};

// Launch with or without the programmatic-stream-serialization attribute (PDL).
template <typename... KArgs, typename... Args>
void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    if (pdl) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    NMF_CUDA(cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...));
}

// Rows of a factor per CTA of the update kernel (multiple of 8, <= 128).
int pick_tile_rows(int R, int forced) {
    if (forced >= 8 && forced <= 128 && forced % 8 == 0) return forced;
    // Measured (profiles/r1b_pipeline_experiments.md): the time per k-block does not shrink with the box height,
    // so full 128-row boxes always win, even when that leaves SMs idle (128 CTAs at 16384 rows).
    return R >= 128 ? 128 : (int)round_up(std::max(R, 8), 8);
}

Factor alloc_factor(nmfb200_handle* h, const std::string& tag, int R, int KP) {
    Factor f;
    std::string t(tag);
    f.R = R;
    f.ldT = round_up(R, 64);
    f.tile_rows = pick_tile_rows(R, h->tc_tile_rows);
    f.tiles = (int)ceil_div(R, f.tile_rows);
    f.m = h->buf_t<float>("tc." + t + ".m", (size_t)R * KP);
    f.hi = h->buf_t<bf16>("tc." + t + ".hi", (size_t)R * KP);
    f.lo = h->buf_t<bf16>("tc." + t + ".lo", (size_t)R * KP);
    f.rowsT = KP < 128 ? 128 : KP;  // gram_kernel loads 128-row M tiles: keep zero rows behind KP = 64
    f.bT = h->buf_t<bf16>("tc." + t + ".bT", (size_t)f.rowsT * f.ldT);
    if (h->tc_precision == 1) {
        const size_t before = h->bufs["tc." + t + ".bTlo"].bytes;
        f.bTlo = h->buf_t<bf16>("tc." + t + ".bTlo", (size_t)f.rowsT * f.ldT);
        if (h->bufs["tc." + t + ".bTlo"].bytes != before)   // fresh allocation: rows behind k (KP = 64: up to 128) must read as zero
            NMF_CUDA(cudaMemsetAsync(f.bTlo, 0, (size_t)f.rowsT * f.ldT * sizeof(bf16), h->stream));
    }
    f.P = h->buf_t<float>("tc." + t + ".P", (size_t)KP * KP);
    f.Phi = h->buf_t<bf16>("tc." + t + ".Phi", (size_t)KP * KP);
    f.Plo = h->buf_t<bf16>("tc." + t + ".Plo", (size_t)KP * KP);
    f.conv = h->buf_t<float>("tc." + t + ".conv", (size_t)f.tiles * 2 * KP);
    f.colsum = h->buf_t<float>("tc." + t + ".colsum", (size_t)KP);
    return f;
}

// ||X||^2 in Float64 (trace-identity objective of the verbose path), once per set_X
__global__ void sumsq_kernel(const float* __restrict__ X, int64_t p, int64_t n, int64_t ldx, double* __restrict__ part) {
    __shared__ double red[256];
    double s = 0.0;
    const int64_t total = p * n;
    for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
        const float v = X[(t % p) + (t / p) * ldx];
        s += (double)v * (double)v;
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}
double x_norm2(nmfb200_handle* h) {
    if (h->x_norm2_epoch == h->x_epoch) return h->x_norm2;
    const int nb = 148 * 8;
    double* part = h->buf_t<double>("tc.xnorm_part", nb);
    sumsq_kernel<<<nb, 256, 0, h->stream>>>((const float*)h->dX, h->p, h->n, h->ldx, part);
    h->launches += 1;
    std::vector<double> hp(nb);
    NMF_CUDA(cudaMemcpyAsync(hp.data(), part, nb * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    NMF_CUDA(cudaStreamSynchronize(h->stream));
    double s = 0.0;
    for (double v : hp) s += v;
    h->x_norm2 = s;
    h->x_norm2_epoch = h->x_epoch;
    return s;
}

// bf16 tile-contiguous caches of X in both orientations (built once per set_X / tile shape / shard geometry).
// X, p, ldx describe the rows this (logical) rank works on: the whole matrix, or a row shard of it (pfx names the buffers).
void build_x_caches(nmfb200_handle* h, const std::string& pfx, const float* X, int64_t p, int64_t n, int64_t ldx, bf16** Xr_out,
                    bf16** Xc_out, bf16** Xr_lo_out = nullptr, bf16** Xc_lo_out = nullptr) {
    cudaStream_t st = h->stream;
    const int trH = pick_tile_rows((int)n, h->tc_tile_rows), trW = pick_tile_rows((int)p, h->tc_tile_rows);
    const int64_t tilesH = ceil_div(n, trH), tilesW = ceil_div(p, trW);
    const int nkbH = (int)ceil_div(p, 64), nkbW = (int)ceil_div(n, 64);
    bf16* Xr_ = h->buf_t<bf16>(pfx + ".Xr", (size_t)tilesH * nkbH * trH * 64);  // rows = columns of X, contraction over p
    bf16* Xc_ = h->buf_t<bf16>(pfx + ".Xc", (size_t)tilesW * nkbW * trW * 64);  // rows = rows of X, contraction over n
    const bool x3 = h->tc_precision == 1 && Xr_lo_out != nullptr;
    bf16* Xr_lo = x3 ? h->buf_t<bf16>(pfx + ".Xr_lo", (size_t)tilesH * nkbH * trH * 64) : nullptr;
    bf16* Xc_lo = x3 ? h->buf_t<bf16>(pfx + ".Xc_lo", (size_t)tilesW * nkbW * trW * 64) : nullptr;
    nmfb200_handle::XCacheKey key{h->x_epoch, (const void*)X, p, trH, trW, x3 ? 1 : 0};
    nmfb200_handle::XCacheKey& have = h->tc_x_cache[pfx];
    if (!have.covers(key)) {
        cvt_tiled_direct_kernel<<<dim3((unsigned)(tilesH * trH), (unsigned)std::min<int64_t>(ceil_div((int64_t)nkbH * 64, 256), 64)), 256, 0, st>>>(
            X, ldx, (int)n, (int)p, trH, nkbH, Xr_, Xr_lo);
        NMF_REQUIRE(trW % 8 == 0, NMFB200_EINVAL, "tile rows must be a multiple of 8");
        // the transpose kernel walks 64 logical rows per block; cover the padded row range of the last tile too
        cvt_tiled_transpose_kernel<<<dim3((unsigned)nkbW, (unsigned)ceil_div(tilesW * trW, 64)), dim3(32, 8), 0, st>>>(
            X, ldx, (int)p, (int)n, trW, nkbW, (int)tilesW, Xc_, Xc_lo);
        h->launches += 2;
        NMF_CUDA(cudaGetLastError());
        have = key;
    }
    *Xr_out = Xr_;
    *Xc_out = Xc_;
    if (Xr_lo_out) *Xr_lo_out = Xr_lo;
    if (Xc_lo_out) *Xc_lo_out = Xc_lo;
}
void build_x_caches(nmfb200_handle* h, bf16** Xr_out, bf16** Xc_out, bf16** Xr_lo_out = nullptr, bf16** Xc_lo_out = nullptr) {
    build_x_caches(h, "tc", (const float*)h->dX, h->p, h->n, h->ldx, Xr_out, Xc_out, Xr_lo_out, Xc_lo_out);
}
