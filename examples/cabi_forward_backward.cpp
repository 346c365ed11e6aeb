// Calling the rasterizer through its C ABI from a compiled host, without PyTorch.
//
// INTEGRATION.md shows the ctypes binding a maintainer of the (Python) reference adds; this is the same
// sequence from C++ with plain CUDA-runtime allocations — what a C++ caller such as upstream's
// rasterize_points.cu shim, or a cgo / JNI binding, would do:
//   surfel_forward_preprocess -> wait for R -> size the binning workspace -> surfel_forward_render
//   -> surfel_backward.
// Build (tests/test_oracle_cpu.py compiles and links it; running needs a GPU):
//   g++ -std=c++17 -I include -I /usr/local/cuda/include examples/cabi_forward_backward.cpp
//       2d-gaussian-splatting_b200/lib/libsurfel_b200.so -L /usr/local/cuda/lib64 -lcudart -o cabi_example
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "surfel_rasterizer.h"

#define CUDA_OK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define SURFEL_OK(x) do { if ((x) != 0) { std::fprintf(stderr, "%s: %s\n", #x, surfel_last_error()); return 3; } } while (0)

template <class T>
static T* to_device(const std::vector<T>& h) {
    T* d = nullptr;
    if (cudaMalloc(&d, h.size() * sizeof(T)) != cudaSuccess) return nullptr;
    cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    return d;
}
template <class T>
static T* device_buffer(size_t n) {
    T* d = nullptr;
    return cudaMalloc(&d, (n ? n : 1) * sizeof(T)) == cudaSuccess ? d : nullptr;
}

int main() {
    if (surfel_abi_version() != SURFEL_ABI_VERSION) { std::fprintf(stderr, "ABI mismatch\n"); return 1; }
    const int P = 2000, M = 16, W = 320, H = 240;
    const float tanfovy = std::tan(25.0f * 3.14159265f / 180.0f), tanfovx = tanfovy * W / H;
    const float zn = 0.01f, zf = 100.0f;
    // camera at the origin looking down +z; matrices in the reference's row-vector convention (p_row @ M)
    std::vector<float> view = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    std::vector<float> proj = {1 / tanfovx, 0, 0, 0, 0, 1 / tanfovy, 0, 0, 0, 0, zf / (zf - zn), 1, 0, 0, -(zf * zn) / (zf - zn), 0};
    std::vector<float> campos = {0, 0, 0}, bg = {0, 0, 0};

    std::srand(7);
    auto u = [] { return std::rand() / (float)RAND_MAX; };
    std::vector<float> means(3 * P), scales(2 * P), rots(4 * P), opac(P), shs((size_t)P * M * 3);
    for (int i = 0; i < P; i++) {
        const float z = 2.0f + 10.0f * u();
        means[3 * i] = (2 * u() - 1) * tanfovx * z; means[3 * i + 1] = (2 * u() - 1) * tanfovy * z; means[3 * i + 2] = z;
        scales[2 * i] = z * (0.5f + 3 * u()) * 2 * tanfovx / W; scales[2 * i + 1] = z * (0.5f + 3 * u()) * 2 * tanfovx / W;
        float q[4] = {2 * u() - 1, 2 * u() - 1, 2 * u() - 1, 2 * u() - 1};
        const float n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]) + 1e-9f;
        for (int k = 0; k < 4; k++) rots[4 * i + k] = q[k] / n;
        opac[i] = 0.1f + 0.8f * u();
        for (int k = 0; k < M * 3; k++) shs[(size_t)i * M * 3 + k] = (k < 3 ? 1.0f : 0.2f) * (2 * u() - 1);
    }

    surfel_settings_t s = {};
    s.image_height = H; s.image_width = W; s.tanfovx = tanfovx; s.tanfovy = tanfovy; s.scale_modifier = 1.0f; s.sh_degree = 3;
    s.bg = to_device(bg); s.viewmatrix = to_device(view); s.projmatrix = to_device(proj); s.campos = to_device(campos);
    float *d_means = to_device(means), *d_scales = to_device(scales), *d_rots = to_device(rots), *d_opac = to_device(opac), *d_shs = to_device(shs);
    int32_t* d_radii = device_buffer<int32_t>(P);
    void* geom = device_buffer<char>(surfel_geom_bytes(P));
    void* image = device_buffer<char>(surfel_image_bytes(W, H));
    float *d_color = device_buffer<float>(3 * (size_t)W * H), *d_others = device_buffer<float>(7 * (size_t)W * H);
    uint32_t* R_host = nullptr;
    CUDA_OK(cudaHostAlloc((void**)&R_host, sizeof(uint32_t), cudaHostAllocDefault));   // pinned + mapped: R arrives by a zero-copy store
    cudaStream_t stream;
    CUDA_OK(cudaStreamCreate(&stream));

    SURFEL_OK(surfel_forward_preprocess(&s, P, M, d_means, d_opac, d_scales, d_rots, nullptr, d_shs, nullptr, d_radii, geom,
                                        image /* per-tile counts fused in */, R_host, stream));
    CUDA_OK(cudaStreamSynchronize(stream));            // the one host wait of the forward (upstream blocks here too)
    const uint32_t R = *R_host;
    void* binning = device_buffer<char>(surfel_binning_bytes(R, W, H));
    SURFEL_OK(surfel_forward_render(&s, P, R, d_radii, geom, binning, image, /*tile_counts_ready=*/1, d_color, d_others, stream));

    // backward with dL/dcolor = 1, dL/dothers = 0
    std::vector<float> ones(3 * (size_t)W * H, 1.0f), zeros(7 * (size_t)W * H, 0.0f);
    float *g_color = to_device(ones), *g_others = to_device(zeros);
    float* scratch = device_buffer<float>((size_t)P * surfel_grad_scratch_floats());
    float *g_m2d = device_buffer<float>(3 * P), *g_opac = device_buffer<float>(P), *g_m3d = device_buffer<float>(3 * P);
    float *g_sh = device_buffer<float>((size_t)P * M * 3), *g_scales = device_buffer<float>(2 * P), *g_rots = device_buffer<float>(4 * P);
    SURFEL_OK(surfel_backward(&s, P, M, R, d_means, d_scales, d_rots, nullptr, d_shs, /*has_colors_precomp=*/0, d_radii, geom, binning,
                              image, g_color, g_others, scratch, g_m2d, /*dL_dcolors=*/nullptr, g_opac, g_m3d,
                              /*dL_dtransMat=*/nullptr, g_sh, g_scales, g_rots, /*lowpass_depth_quirk=*/0, stream));
    std::vector<float> color(3 * (size_t)W * H), gop(P);
    CUDA_OK(cudaMemcpyAsync(color.data(), d_color, color.size() * sizeof(float), cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaMemcpyAsync(gop.data(), g_opac, gop.size() * sizeof(float), cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
    double sum = 0, gsum = 0;
    for (float v : color) sum += v;
    for (float v : gop) gsum += v;
    std::printf("P=%d  R=%u  mean colour=%.6f  sum dL/dopacity=%.6f\n", P, R, sum / color.size(), gsum);
    return 0;
}
