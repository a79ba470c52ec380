/*
 * A plain-C host of the C-ABI (include/gsvc_rast.h): no Python, no torch — device memory from the CUDA runtime,
 * the allocator-callback forward (the shape of the upstream binding, three opaque buffers resized through a
 * callback) and the backward, the way a compiled plugin such as the reference's `_C.rasterize_gaussians` /
 * `_C.rasterize_gaussians_backward` pair (diff_gaussian_rasterization/__init__.py of the un-vendored dependency,
 * called from ortho_gaussian_renderer/renderer.py:100-109) would sit on top of libgsvc_rast.so.
 *
 *   gcc -std=c99 -O1 -Iinclude -I/usr/local/cuda/include examples/c_host.c -o c_host \
 *       -Lgsvc_b200 -lgsvc_rast -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/gsvc_b200 -Wl,-rpath,/usr/local/cuda/lib64
 *   ./c_host scene.bin out.bin
 *
 * scene.bin (little endian): int32 W, H, P; float x_min, y_min, scale, threshold, scale_modifier, bg[3], V[16]
 * (logical, row-major); float means3D[P*3], opacities[P], colors[P*3], scales[P*3], rotations[P*4], dL_dout[3*H*W].
 * out.bin: int64 num_rendered; float color[3*H*W]; int32 radii[P]; float dL_dmeans3D[P*3], dL_dcolors[P*3],
 * dL_dopacities[P], dL_dscales[P*3], dL_drotations[P*4].
 * tests/test_gpu_parity.py::test_plain_c_host_matches_oracle builds it, runs it and checks out.bin against the oracle.
 */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gsvc_rast.h"

#define CK(x)                                                                                 \
    do {                                                                                      \
        cudaError_t e_ = (x);                                                                 \
        if (e_ != cudaSuccess) {                                                              \
            fprintf(stderr, "%s:%d: %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(2);                                                                          \
        }                                                                                     \
    } while (0)

/* the three opaque state buffers of a forward, grown on demand (which: 0 geom, 1 binning, 2 image) */
typedef struct {
    void *ptr[3];
    size_t cap[3];
} state_buffers;

static void *grow(void *user, int32_t which, size_t bytes)
{
    state_buffers *s = (state_buffers *)user;
    if (which < 0 || which > 2) return NULL;
    if (bytes > s->cap[which]) {
        if (s->ptr[which]) cudaFree(s->ptr[which]);
        if (cudaMalloc(&s->ptr[which], bytes) != cudaSuccess) return NULL;
        s->cap[which] = bytes;
    }
    return s->ptr[which];
}

static float *upload(FILE *f, size_t n)
{
    float *h = (float *)malloc(n * sizeof(float) + 4), *d = NULL;
    if (fread(h, sizeof(float), n, f) != n) {
        fprintf(stderr, "scene file too short\n");
        exit(2);
    }
    CK(cudaMalloc((void **)&d, n * sizeof(float) + 4));
    CK(cudaMemcpy(d, h, n * sizeof(float), cudaMemcpyHostToDevice));
    free(h);
    return d;
}

static void download(FILE *f, const void *d, size_t bytes)
{
    void *h = malloc(bytes + 4);
    CK(cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost));
    fwrite(h, 1, bytes, f);
    free(h);
}

int main(int argc, char **argv)
{
    if (argc != 3) {
        fprintf(stderr, "usage: %s scene.bin out.bin\n", argv[0]);
        return 2;
    }
    if (gsvc_rast_abi_version() != GSVC_RAST_ABI_VERSION) {
        fprintf(stderr, "libgsvc_rast.so ABI %d, header %d\n", gsvc_rast_abi_version(), GSVC_RAST_ABI_VERSION);
        return 2;
    }
    FILE *f = fopen(argv[1], "rb");
    if (!f) {
        perror(argv[1]);
        return 2;
    }
    int32_t dims[3];
    float hdr[5 + 3 + 16];
    if (fread(dims, 4, 3, f) != 3 || fread(hdr, 4, 24, f) != 24) return 2;
    const int32_t W = dims[0], H = dims[1], P = dims[2];
    const size_t N = (size_t)W * H;

    cudaStream_t stream;
    CK(cudaStreamCreate(&stream));
    float *bg, *V;
    CK(cudaMalloc((void **)&bg, 3 * sizeof(float)));
    CK(cudaMalloc((void **)&V, 16 * sizeof(float)));
    CK(cudaMemcpy(bg, hdr + 5, 3 * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(V, hdr + 8, 16 * sizeof(float), cudaMemcpyHostToDevice));
    float *means = upload(f, (size_t)P * 3), *opac = upload(f, P), *colors = upload(f, (size_t)P * 3);
    float *scales = upload(f, (size_t)P * 3), *rots = upload(f, (size_t)P * 4), *dL = upload(f, 3 * N);
    fclose(f);

    gsvc_rast_settings st;
    memset(&st, 0, sizeof st);
    st.image_width = W;
    st.image_height = H;
    st.x_min = hdr[0];
    st.y_min = hdr[1];
    st.scale = hdr[2];
    st.threshold = hdr[3];
    st.scale_modifier = hdr[4];
    st.bg = bg;
    st.viewmatrix = V; /* a contiguous logical matrix: row stride 4, column stride 1 */
    st.vm_stride_r = 4;
    st.vm_stride_c = 1;

    float *color, *g_means, *g_colors, *g_opac, *g_scales, *g_rots;
    int32_t *radii;
    void *scratch;
    CK(cudaMalloc((void **)&color, 3 * N * sizeof(float)));
    CK(cudaMalloc((void **)&radii, (size_t)P * sizeof(int32_t) + 4));
    CK(cudaMalloc((void **)&g_means, (size_t)P * 3 * sizeof(float) + 4));
    CK(cudaMalloc((void **)&g_colors, (size_t)P * 3 * sizeof(float) + 4));
    CK(cudaMalloc((void **)&g_opac, (size_t)P * sizeof(float) + 4));
    CK(cudaMalloc((void **)&g_scales, (size_t)P * 3 * sizeof(float) + 4));
    CK(cudaMalloc((void **)&g_rots, (size_t)P * 4 * sizeof(float) + 4));
    CK(cudaMalloc(&scratch, gsvc_rast_backward_scratch_bytes(P) + 4));

    state_buffers state;
    memset(&state, 0, sizeof state);
    const int64_t R = gsvc_rast_forward(&st, P, 0, means, NULL, colors, opac, scales, rots, NULL, grow, &state, color,
                                        radii, stream);
    if (R < 0) {
        fprintf(stderr, "gsvc_rast_forward: %lld: %s\n", (long long)R, gsvc_rast_last_error());
        return 1;
    }
    const int rc = gsvc_rast_backward(&st, P, 0, R > 0 ? R : 1, means, NULL, colors, scales, rots, NULL, radii,
                                      state.ptr[0], state.ptr[2], state.ptr[1], scratch, /*scratch_is_zero=*/0, dL,
                                      g_means, NULL, g_colors, g_opac, g_scales, g_rots, NULL, NULL, NULL, stream);
    if (rc != GSVC_RAST_OK) {
        fprintf(stderr, "gsvc_rast_backward: %d: %s\n", rc, gsvc_rast_last_error());
        return 1;
    }
    CK(cudaStreamSynchronize(stream));

    FILE *o = fopen(argv[2], "wb");
    if (!o) {
        perror(argv[2]);
        return 2;
    }
    fwrite(&R, sizeof R, 1, o);
    download(o, color, 3 * N * sizeof(float));
    download(o, radii, (size_t)P * sizeof(int32_t));
    download(o, g_means, (size_t)P * 3 * sizeof(float));
    download(o, g_colors, (size_t)P * 3 * sizeof(float));
    download(o, g_opac, (size_t)P * sizeof(float));
    download(o, g_scales, (size_t)P * 3 * sizeof(float));
    download(o, g_rots, (size_t)P * 4 * sizeof(float));
    fclose(o);
    printf("num_rendered %lld, %lld kernel launches\n", (long long)R, (long long)gsvc_rast_launch_count(0));
    return 0;
}
