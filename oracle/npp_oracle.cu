// npp_oracle.cu -- TEST INFRASTRUCTURE (see oracle/__init__.py): the reference's NPP letter-box branch, executed by the same
// NPP call the reference makes, so that the product's area-resampling kernel can be compared with it on a GPU box.
// Restates RetinaFace::detect's USE_NPP preprocess (retinaface/RetinaFace.cpp:594-600: H2D of the BGR u8 image, memset of the
// network-sized buffer, imageROIResize8U3C over the whole image) and imageROIResize8U3C itself
// (retinaface/resizeconvertion.cu:279-316: one isotropic factor = min(dstW/srcW, dstH/srcH) clamped to <= 1, shifts 0,
// NPPI_INTER_SUPER when shrinking, NPPI_INTER_LANCZOS at factor 1, destination anchored top-left).
// NPP is closed source (CUDA toolkit, libnppig 12.4 in this image); nothing of it ships in the product.
//   nvcc -shared -Xcompiler -fPIC -o oracle/libnpp_oracle.so oracle/npp_oracle.cu -lnppig -lnppc
#include <cuda_runtime.h>
#include <nppi_geometry_transforms.h>
#include <stdint.h>

extern "C" int npp_letterbox(const uint8_t *host_bgr, int w, int h, uint8_t *host_out, int net_w, int net_h) {
    uint8_t *d_src = nullptr, *d_dst = nullptr;
    if (cudaMalloc(&d_src, (size_t)w * h * 3) != cudaSuccess) return -1;
    if (cudaMalloc(&d_dst, (size_t)net_w * net_h * 3) != cudaSuccess) { cudaFree(d_src); return -1; }
    cudaMemcpy(d_src, host_bgr, (size_t)w * h * 3, cudaMemcpyHostToDevice);       // RetinaFace.cpp:594
    cudaMemset(d_dst, 0, (size_t)net_w * net_h * 3);                                // RetinaFace.cpp:598
    NppiSize src_size = {w, h};
    NppiRect src_roi = {0, 0, w, h};                                                // RetinaFace.cpp:599: whole image
    NppiRect dst_roi = {0, 0, net_w, net_h};
    const double fx = (double)net_w / w, fy = (double)net_h / h;
    double f = fx < fy ? fx : fy;
    if (f > 1.0) f = 1.0;                                                           // never up-scaled
    const int interp = f >= 1.0 ? NPPI_INTER_LANCZOS : NPPI_INTER_SUPER;
    NppStatus st = nppiResizeSqrPixel_8u_C3R(d_src, src_size, w * 3, src_roi, d_dst, net_w * 3, dst_roi, f, f, 0.0, 0.0, interp);
    cudaMemcpy(host_out, d_dst, (size_t)net_w * net_h * 3, cudaMemcpyDeviceToHost);
    cudaFree(d_src);
    cudaFree(d_dst);
    return (int)st;
}
