// tex_floor.cu -- which texel does point sampling with unnormalised coordinates return next to a texel boundary?
// The line probes of k_forward_line / k_reverse may read the distance bytes through a 3-D texture (DMF_LINE_TEX); their
// safety argument needs a bound tau on |coordinate the texture unit floors - coordinate passed|.  For every axis and a set
// of integers i this prints the largest j for which x = i - 2^-j still returns texel i-1 (i.e. is floored correctly), and
// the same from above (x = i + 2^-j -> texel i is trivially right; what matters is rounding UP across a boundary).
//   nvcc -arch=sm_100a -o /tmp/tex_floor tools/micro/tex_floor.cu && /tmp/tex_floor
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

__global__ void k_probe(cudaTextureObject_t t, const float* xyz, unsigned* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = tex3D<unsigned char>(t, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}

int main() {
    const int NX = 521, NY = 37, NZ = 29;       // tex x, y, z
    std::vector<unsigned char> h((size_t)NX * NY * NZ);
    // texel value: low 3 bits of each coordinate -> neighbours differ on every axis
    for (int z = 0; z < NZ; z++) for (int y = 0; y < NY; y++) for (int x = 0; x < NX; x++) h[((size_t)z * NY + y) * NX + x] = (unsigned char)((x & 7) | ((y & 7) << 3) | ((z & 3) << 6));
    cudaArray_t arr; cudaChannelFormatDesc desc = cudaCreateChannelDesc(8, 0, 0, 0, cudaChannelFormatKindUnsigned);
    cudaExtent ext = make_cudaExtent(NX, NY, NZ);
    if (cudaMalloc3DArray(&arr, &desc, ext) != cudaSuccess) { printf("cudaMalloc3DArray failed\n"); return 1; }
    cudaMemcpy3DParms cp; memset(&cp, 0, sizeof cp);
    cp.srcPtr = make_cudaPitchedPtr(h.data(), NX, NX, NY); cp.dstArray = arr; cp.extent = ext; cp.kind = cudaMemcpyHostToDevice;
    cudaMemcpy3D(&cp);
    cudaResourceDesc rd; memset(&rd, 0, sizeof rd); rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
    cudaTextureDesc td; memset(&td, 0, sizeof td);
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex; cudaCreateTextureObject(&tex, &rd, &td, nullptr);

    std::vector<float> q; std::vector<int> exp_;
    struct Case { int axis, i, j, below; }; std::vector<Case> cases;
    const int dims[3] = {NX, NY, NZ};
    for (int axis = 0; axis < 3; axis++)
        for (int i : {1, 2, 3, 5, 8, 16, 17, 28, 100, 255, 256, 511, 520}) {
            if (i >= dims[axis]) continue;
            for (int j = 1; j <= 23; j++)
                for (int below = 0; below < 2; below++) {
                    float c[3] = {3.5f, 4.5f, 2.5f};
                    const float x = below ? (float)i - ldexpf(1.f, -j) : (float)i + ldexpf(1.f, -j);
                    if (x == (float)i) continue;                              // not representable at this magnitude
                    c[axis] = x;
                    int t[3] = {3, 4, 2}; t[axis] = below ? i - 1 : i;
                    q.push_back(c[0]); q.push_back(c[1]); q.push_back(c[2]);
                    exp_.push_back((t[0] & 7) | ((t[1] & 7) << 3) | ((t[2] & 3) << 6));
                    cases.push_back({axis, i, j, below});
                }
        }
    // random points well inside texels (fraction in [2^-6, 1 - 2^-6]) over the whole array: must all be exact
    unsigned long long s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0; };
    const size_t n_edge = cases.size();
    for (int r = 0; r < 200000; r++) {
        int t[3]; float c[3];
        for (int a = 0; a < 3; a++) { t[a] = (int)(rnd() * dims[a]); c[a] = (float)(t[a] + 1.0 / 64 + rnd() * (1.0 - 2.0 / 64)); }
        q.push_back(c[0]); q.push_back(c[1]); q.push_back(c[2]);
        exp_.push_back((t[0] & 7) | ((t[1] & 7) << 3) | ((t[2] & 3) << 6));
    }
    const int n = (int)exp_.size();
    float* dq; unsigned* dout; cudaMalloc(&dq, q.size() * 4); cudaMalloc(&dout, n * 4);
    cudaMemcpy(dq, q.data(), q.size() * 4, cudaMemcpyHostToDevice);
    k_probe<<<(n + 255) / 256, 256>>>(tex, dq, dout, n);
    std::vector<unsigned> out(n); cudaMemcpy(out.data(), dout, n * 4, cudaMemcpyDeviceToHost);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed\n"); return 1; }
    // per (axis, below): the smallest j (= largest offset 2^-j) that was floored WRONGLY
    int worst[3][2]; for (auto& w : worst) w[0] = w[1] = 99;
    for (size_t k = 0; k < n_edge; k++)
        if ((int)out[k] != exp_[k]) { auto& w = worst[cases[k].axis][cases[k].below]; if (cases[k].j < w) w = cases[k].j; }
    for (int a = 0; a < 3; a++)
        printf("axis %d: i - 2^-j wrong from j = %d on;  i + 2^-j wrong from j = %d on (99 = never)\n", a, worst[a][1], worst[a][0]);
    size_t bad = 0; for (size_t k = n_edge; k < (size_t)n; k++) bad += (int)out[k] != exp_[k];
    printf("random interior points (fraction in [1/64, 63/64]): %zu of %zu wrong\n", bad, (size_t)n - n_edge);
    return 0;
}
