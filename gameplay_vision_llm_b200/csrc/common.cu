// common.cu — host-side plumbing of libgvl_sm100a.so: thread-local error string, launch counter,
// device checks and the TMA descriptor helper.
#include "common.cuh"

#include <cstdlib>

#include <mutex>
#include <vector>

namespace gvl {

static thread_local char t_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("GVL_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// ---- per-launch profiling ----
struct ProfRec {
    int kid;
    double work;
    cudaEvent_t e0, e1;
};
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRec*> g_prof_recs;   // records of the current session
static std::vector<ProfRec*> g_prof_pool;   // recycled records (events are reused)

ProfScope::ProfScope(int kernel_id, double work, cudaStream_t s) : rec(nullptr), stream(s) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec* r = nullptr;
    if (!g_prof_pool.empty()) {
        r = g_prof_pool.back();
        g_prof_pool.pop_back();
    } else {
        r = new ProfRec();
        if (cudaEventCreate(&r->e0) != cudaSuccess || cudaEventCreate(&r->e1) != cudaSuccess) {
            delete r;
            return;
        }
    }
    r->kid = kernel_id;
    r->work = work;
    cudaEventRecord(r->e0, s);
    g_prof_recs.push_back(r);
    rec = r;
}
ProfScope::~ProfScope() {
    if (rec) cudaEventRecord(static_cast<ProfRec*>(rec)->e1, stream);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols, bool swizzle128) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
        return 4;
    }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estride[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d): base=%p rows=%llu cols=%llu ld=%llu box=%ux%u", (int)r,
                  base, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
        return 4;
    }
    return 0;
}

int make_tmap_nd_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, int swizzle_bytes) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
        return 4;
    }
    cuuint64_t gdim[5], gstride[4];
    cuuint32_t bx[5], estride[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        estride[i] = 1;
        if (i > 0) gstride[i - 1] = strides_bytes[i - 1];
    }
    CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                            : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                            : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                  : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstride, bx,
                    estride, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (rank %d) failed (CUresult %d)", rank, (int)r);
        return 4;
    }
    return 0;
}

}  // namespace gvl

extern "C" {

const char* gvl_last_error(void) { return gvl::t_err; }
int gvl_abi_version(void) { return GVL_ABI_VERSION; }
unsigned long long gvl_launch_count(void) { return gvl::g_launches.load(); }

int gvl_prof_enable(int on) {
    std::lock_guard<std::mutex> lk(gvl::g_prof_mu);
    if (on) {  // a new session recycles the previous one's records; disabling keeps them readable
        for (gvl::ProfRec* r : gvl::g_prof_recs) gvl::g_prof_pool.push_back(r);
        gvl::g_prof_recs.clear();
    }
    gvl::g_prof_on = on != 0;
    return 0;
}

int gvl_prof_summary(int kernel_id, double* total_ms, unsigned long long* launches, double* total_work) {
    std::lock_guard<std::mutex> lk(gvl::g_prof_mu);
    double ms = 0.0, work = 0.0;
    unsigned long long n = 0;
    for (gvl::ProfRec* r : gvl::g_prof_recs) {
        if (r->kid != kernel_id) continue;
        GVL_CUDA(cudaEventSynchronize(r->e1));
        float t = 0.f;
        GVL_CUDA(cudaEventElapsedTime(&t, r->e0, r->e1));
        ms += t;
        work += r->work;
        ++n;
    }
    if (total_ms) *total_ms = ms;
    if (launches) *launches = n;
    if (total_work) *total_work = work;
    return 0;
}

int gvl_check_device(int dev) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        gvl::set_error("no CUDA device: %s", cudaGetErrorString(e));
        return 2;
    }
    cudaDeviceProp prop;
    GVL_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        gvl::set_error("device %d is sm_%d%d; libgvl_sm100a needs sm_100 (B200)", dev, prop.major, prop.minor);
        return 5;
    }
    return 0;
}

}  // extern "C"
