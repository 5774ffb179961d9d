// JPEG file bytes -> uint8 HWC BGR image in device memory (image-file ingest).
// Replaces the cv2.imread of LoadImagesAndVideos.__next__ (ultralytics/data/loaders.py:284-448, imread at :406) for
// JPEG sources: the Huffman stage runs on the host inside nvJPEG, the IDCT / upsampling / colour conversion on the GPU,
// and the decoded pixels are born in HBM in the layout specyolo_letterbox_u8 consumes (cv2's HWC BGR) — no host image, no
// H2D copy of raw pixels.  nvJPEG is a LIBRARY call (like cuBLAS); it is loaded with dlopen at first use so that the
// rest of libspecyolo never depends on it.
#include <dlfcn.h>
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <vector>

#include <nvjpeg.h>

#include "common.h"

namespace specyolo {

namespace {
struct NvJpeg {
    void* lib = nullptr;
    nvjpegHandle_t handle = nullptr;
    static constexpr int kStates = 16;      // decode states: one per concurrently decoding host thread
    nvjpegJpegState_t state[kStates] = {};
    std::mutex state_mu[kStates];
    nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*) = nullptr;
    nvjpegStatus_t (*CreateEx)(nvjpegBackend_t, nvjpegDevAllocator_t*, nvjpegPinnedAllocator_t*, unsigned int, nvjpegHandle_t*) = nullptr;
    nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
    nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
    nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t,
                             nvjpegImage_t*, cudaStream_t) = nullptr;
    nvjpegStatus_t (*BatchedInit)(nvjpegHandle_t, nvjpegJpegState_t, int, int, nvjpegOutputFormat_t) = nullptr;
    nvjpegStatus_t (*DecodeBatched)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char* const*, const size_t*, nvjpegImage_t*,
                                    cudaStream_t) = nullptr;
    // batched decode (GPU-assisted Huffman or the hardware engine): one handle + state per back end, created on demand
    nvjpegHandle_t b_handle[4] = {};
    nvjpegJpegState_t b_state[4] = {};
    int b_size[4] = {0, 0, 0, 0};
    std::mutex b_mu;
    std::mutex mu;          // guards initialisation; decodes take one of the per-state mutexes (the Huffman stage of
                            // nvjpegDecode runs on the calling host thread: a loader thread pool scales it with the cores)
    int status = -1;        // -1 not tried, 0 ready, > 0 failed
};
NvJpeg g_nj;

int ensure_nvjpeg() {
    if (g_nj.status >= 0) return g_nj.status;
    const char* names[] = {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"};
    for (const char* n : names) {
        g_nj.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (g_nj.lib) break;
    }
    if (!g_nj.lib) return g_nj.status = 1;
    g_nj.CreateSimple = reinterpret_cast<decltype(g_nj.CreateSimple)>(dlsym(g_nj.lib, "nvjpegCreateSimple"));
    g_nj.JpegStateCreate = reinterpret_cast<decltype(g_nj.JpegStateCreate)>(dlsym(g_nj.lib, "nvjpegJpegStateCreate"));
    g_nj.GetImageInfo = reinterpret_cast<decltype(g_nj.GetImageInfo)>(dlsym(g_nj.lib, "nvjpegGetImageInfo"));
    g_nj.Decode = reinterpret_cast<decltype(g_nj.Decode)>(dlsym(g_nj.lib, "nvjpegDecode"));
    if (!g_nj.CreateSimple || !g_nj.JpegStateCreate || !g_nj.GetImageInfo || !g_nj.Decode) return g_nj.status = 2;
    g_nj.CreateEx = reinterpret_cast<decltype(g_nj.CreateEx)>(dlsym(g_nj.lib, "nvjpegCreateEx"));
    g_nj.BatchedInit = reinterpret_cast<decltype(g_nj.BatchedInit)>(dlsym(g_nj.lib, "nvjpegDecodeBatchedInitialize"));
    g_nj.DecodeBatched = reinterpret_cast<decltype(g_nj.DecodeBatched)>(dlsym(g_nj.lib, "nvjpegDecodeBatched"));
    // SPECYOLO_NVJPEG_BACKEND = 1 (hybrid: CPU Huffman) | 2 (GPU-assisted Huffman) | 3 (hardware engine); default: the library's
    const char* be = std::getenv("SPECYOLO_NVJPEG_BACKEND");
    if (be && be[0] >= '1' && be[0] <= '3' && g_nj.CreateEx) {
        if (g_nj.CreateEx(static_cast<nvjpegBackend_t>(be[0] - '0'), nullptr, nullptr, 0, &g_nj.handle) != NVJPEG_STATUS_SUCCESS)
            return g_nj.status = 3;
    } else if (g_nj.CreateSimple(&g_nj.handle) != NVJPEG_STATUS_SUCCESS) {
        return g_nj.status = 3;
    }
    for (int i = 0; i < NvJpeg::kStates; ++i)
        if (g_nj.JpegStateCreate(g_nj.handle, &g_nj.state[i]) != NVJPEG_STATUS_SUCCESS) return g_nj.status = 4;
    return g_nj.status = 0;
}
}  // namespace

int jpeg_info(const void* data, size_t nbytes, int* H, int* W, int* channels) {
    std::lock_guard<std::mutex> lk(g_nj.mu);
    SY_CHECK(ensure_nvjpeg() == 0, SPECYOLO_ERR_UNSUPPORTED, "nvJPEG is not available (dlopen / init step %d)", g_nj.status);
    int nc = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t ss;
    const nvjpegStatus_t st = g_nj.GetImageInfo(g_nj.handle, static_cast<const unsigned char*>(data), nbytes, &nc, &ss, ws, hs);
    SY_CHECK(st == NVJPEG_STATUS_SUCCESS, SPECYOLO_ERR_INVALID, "not a decodable JPEG stream (nvjpeg status %d)", (int)st);
    *H = hs[0]; *W = ws[0]; *channels = nc;
    return SPECYOLO_OK;
}

int jpeg_decode_bgr(const void* data, size_t nbytes, void* out_dev, int H, int W, cudaStream_t stream) {
    {
        std::lock_guard<std::mutex> lk(g_nj.mu);
        SY_CHECK(ensure_nvjpeg() == 0, SPECYOLO_ERR_UNSUPPORTED, "nvJPEG is not available (dlopen / init step %d)", g_nj.status);
    }
    int nc = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t ss;
    nvjpegStatus_t st = g_nj.GetImageInfo(g_nj.handle, static_cast<const unsigned char*>(data), nbytes, &nc, &ss, ws, hs);
    SY_CHECK(st == NVJPEG_STATUS_SUCCESS, SPECYOLO_ERR_INVALID, "not a decodable JPEG stream (nvjpeg status %d)", (int)st);
    SY_CHECK(hs[0] == H && ws[0] == W, SPECYOLO_ERR_INVALID, "jpeg is %dx%d, the output buffer was sized for %dx%d", hs[0], ws[0], H, W);
    nvjpegImage_t dst{};
    dst.channel[0] = static_cast<unsigned char*>(out_dev);
    dst.pitch[0] = (size_t)W * 3;
    // a free decode state: first one whose mutex can be taken, else wait for a slot picked by the thread's identity
    static std::atomic<unsigned> next{0};
    int slot = -1;
    for (int i = 0; i < NvJpeg::kStates && slot < 0; ++i) {
        const int c = (int)((next.load(std::memory_order_relaxed) + (unsigned)i) % NvJpeg::kStates);
        if (g_nj.state_mu[c].try_lock()) slot = c;
    }
    if (slot < 0) {
        slot = (int)(next.fetch_add(1, std::memory_order_relaxed) % NvJpeg::kStates);
        g_nj.state_mu[slot].lock();
    } else {
        next.store((unsigned)slot + 1u, std::memory_order_relaxed);
    }
    // grayscale files come out replicated to three channels, as cv2.imread(IMREAD_COLOR) returns them
    st = g_nj.Decode(g_nj.handle, g_nj.state[slot], static_cast<const unsigned char*>(data), nbytes, NVJPEG_OUTPUT_BGRI, &dst, stream);
    g_nj.state_mu[slot].unlock();
    SY_CHECK(st == NVJPEG_STATUS_SUCCESS, SPECYOLO_ERR_CUDA, "nvjpegDecode failed (status %d)", (int)st);
    return SPECYOLO_OK;
}

// n JPEG streams in one nvjpegDecodeBatched call on back end 2 (GPU-assisted Huffman) or 3 (hardware engine).
int jpeg_decode_batch_bgr(const void* const* data, const size_t* nbytes, void* const* out_dev, const int* W, int n, int backend,
                          cudaStream_t stream) {
    {
        std::lock_guard<std::mutex> lk(g_nj.mu);
        SY_CHECK(ensure_nvjpeg() == 0, SPECYOLO_ERR_UNSUPPORTED, "nvJPEG is not available (dlopen / init step %d)", g_nj.status);
    }
    SY_CHECK(backend == 2 || backend == 3, SPECYOLO_ERR_INVALID, "jpeg batch decode: backend must be 2 or 3");
    SY_CHECK(g_nj.CreateEx && g_nj.BatchedInit && g_nj.DecodeBatched, SPECYOLO_ERR_UNSUPPORTED, "nvJPEG batched API not found");
    std::lock_guard<std::mutex> lk(g_nj.b_mu);
    if (!g_nj.b_handle[backend]) {
        nvjpegHandle_t h = nullptr;
        nvjpegStatus_t st = g_nj.CreateEx(static_cast<nvjpegBackend_t>(backend), nullptr, nullptr, 0, &h);
        SY_CHECK(st == NVJPEG_STATUS_SUCCESS, SPECYOLO_ERR_UNSUPPORTED, "nvjpegCreateEx(backend %d) failed (status %d)", backend, (int)st);
        nvjpegJpegState_t s2 = nullptr;
        st = g_nj.JpegStateCreate(h, &s2);
        SY_CHECK(st == NVJPEG_STATUS_SUCCESS, SPECYOLO_ERR_CUDA, "nvjpegJpegStateCreate failed (status %d)", (int)st);
        g_nj.b_handle[backend] = h;
        g_nj.b_state[backend] = s2;
    }
    if (g_nj.b_size[backend] != n) {
        const nvjpegStatus_t st = g_nj.BatchedInit(g_nj.b_handle[backend], g_nj.b_state[backend], n, 1, NVJPEG_OUTPUT_BGRI);
        SY_CHECK(st == NVJPEG_STATUS_SUCCESS, SPECYOLO_ERR_UNSUPPORTED, "nvjpegDecodeBatchedInitialize failed (status %d)", (int)st);
        g_nj.b_size[backend] = n;
    }
    std::vector<nvjpegImage_t> dst((size_t)n);
    for (int i = 0; i < n; ++i) {
        dst[i] = nvjpegImage_t{};
        dst[i].channel[0] = static_cast<unsigned char*>(out_dev[i]);
        dst[i].pitch[0] = (size_t)W[i] * 3;
    }
    const nvjpegStatus_t st = g_nj.DecodeBatched(g_nj.b_handle[backend], g_nj.b_state[backend],
                                                 reinterpret_cast<const unsigned char* const*>(data), nbytes, dst.data(), stream);
    SY_CHECK(st == NVJPEG_STATUS_SUCCESS, SPECYOLO_ERR_UNSUPPORTED, "nvjpegDecodeBatched failed (status %d)", (int)st);
    return SPECYOLO_OK;
}

}  // namespace specyolo
