// Camera -> ray batch on the GPU: replaces the numpy ray generation of the reference's
// eval loader for one pinhole camera -- camera_utils.pixels_to_rays (internal/camera_utils.py:L448-557, perspective,
// no distortion, no NDC) + cast_pinhole_rays / Dataset._make_ray_batch (camera_utils.py:L611-632, datasets.py:L386-476:
// near/far broadcast, cam_dirs = -camtoworld[:3,2], float32 cast at L476) -- so a frame is rendered from
// ~200 bytes of camera parameters instead of a 72-byte-per-ray host batch.
//
// The reference evaluates everything in float64 (int pixel + 0.5 promotes) on float32-valued matrices and casts
// the results to float32; the kernel does the same with explicitly rounded fp64 operations in the same
// association, so the float32 results are identical (tests/test_ray_gen.py pins them against vectors produced by
// the reference's own pixels_to_rays).
#include "../../include/ucnerf_b200.h"
#include "ray_march.cuh"

namespace ucnerf {

__global__ void __launch_bounds__(256)
generate_rays_kernel(const __grid_constant__ CameraConst cam, uint32_t row0, uint32_t n_pix, RayOutPtrs o) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    const uint32_t y = row0 + i / cam.width, x = i % cam.width;
    PixelRay pr;
    pixel_to_ray(cam, (int)x, (int)y, pr);
    if (o.origins) { o.origins[3 * i] = cam.origin[0]; o.origins[3 * i + 1] = cam.origin[1]; o.origins[3 * i + 2] = cam.origin[2]; }
    if (o.cam_dirs) { o.cam_dirs[3 * i] = cam.cam_dir[0]; o.cam_dirs[3 * i + 1] = cam.cam_dir[1]; o.cam_dirs[3 * i + 2] = cam.cam_dir[2]; }
    o.directions[3 * i] = pr.dir[0]; o.directions[3 * i + 1] = pr.dir[1]; o.directions[3 * i + 2] = pr.dir[2];
    o.viewdirs[3 * i] = pr.view[0]; o.viewdirs[3 * i + 1] = pr.view[1]; o.viewdirs[3 * i + 2] = pr.view[2];
    o.radii[i] = pr.radius;
    if (o.near) o.near[i] = cam.near;
    if (o.far) o.far[i] = cam.far;
    if (o.imageplane) { o.imageplane[2 * i] = pr.plane[0]; o.imageplane[2 * i + 1] = pr.plane[1]; }
    if (o.rand_vec) {
        // the cone-basis vector the reference draws with torch.randn_like(cam_dirs) at every call (render.py:L140):
        // any standard-normal draw is a valid realisation; counter-based so it is reproducible per (seed, pixel)
        float n[4];
        normal4(cam.rand_seed, (uint64_t)y * cam.width + x, n);
        o.rand_vec[3 * i] = n[0]; o.rand_vec[3 * i + 1] = n[1]; o.rand_vec[3 * i + 2] = n[2];
    }
}

int launch_generate_rays(const CameraConst& cam, uint32_t row0, uint32_t n_rows, const RayOutPtrs& o, cudaStream_t st) {
    const uint64_t n = (uint64_t)n_rows * cam.width;
    if (n == 0) return 0;
    UC_REQUIRE(n < (1ull << 32), "generate_rays: too many pixels in one call");
    generate_rays_kernel<<<(unsigned)div_up(n, (uint64_t)256), 256, 0, st>>>(cam, row0, (uint32_t)n, o);
    UC_LAUNCH_CHECK();
    return 0;
}

}  // namespace ucnerf
