/*
 * oracle/gut_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, fp32) of the reference 3DGUT render path
 * (nv-tlabs/3dgrut @ a37ef72, threedgut_tracer/).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this; the product
 * (3dgrut_b200/) never does.
 *
 * Pinning status: the reference ships no test or golden vector for this path
 *  and its GPU build cannot run here, so end-to-end parity is
 * "unpinned" in the strict sense.  What IS pinned: every function below that has a
 * hand-written CUDA counterpart in the reference is checked against that counterpart
 * compiled for the host from the reference sources where they lie (oracle/_ref,
 * oracle/ref_gut.cpp, tests/test_oracle_vs_ref.py) and against committed golden
 * vectors generated from it (tests/golden/).
 */
#ifndef GUT_ORACLE_H
#define GUT_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* OpenCV pinhole (model 0), OpenCV fisheye (model 1: radial[0..3] = k1..k4, max_angle) or f-theta (model 2) camera, global shutter
 * (reference: sensors/cameraModels.h:22-35,59-72). */
typedef struct {
    int32_t width, height;
    float principal[2];
    float focal[2];
    float radial[6];
    float tangential[2];
    float thin_prism[4];
    float pose_start[7]; /* t.xyz, q.xyzw ; world -> sensor (sensors.h:33) */
    float pose_end[7];
    int32_t model;       /* TSensorModel::ModelType: 0 OpenCVPinholeModel, 1 OpenCVFisheyeModel, 2 FThetaModel */
    float max_angle;     /* OpenCVFisheyeProjectionParameters::maxAngle / FThetaProjectionParameters::maxAngle */
    /* FThetaProjectionParameters (sensors/cameraModels.h:37-47); principal point = principal[] */
    int32_t ftheta_reference_poly; /* 0 PIXELDIST_TO_ANGLE, 1 ANGLE_TO_PIXELDIST */
    float ftheta_bw[6];            /* pixeldistToAnglePoly (backward) */
    float ftheta_fw[6];            /* angleToPixeldistPoly (forward)  */
    float ftheta_cde[3];           /* linear_cde */
    int32_t rolling_shutter;       /* 0 global shutter, 1..4 = TSensorModel::ShutterType + 1: rolling top-to-bottom, left-to-right,
                                    * bottom-to-top, right-to-left (sensors/cameraModels.h:49-57) */
} gut_oracle_camera;

/* Render configuration = the reference's compile-time -D constants (setup_3dgut.py:64-95). */
typedef struct {
    int32_t kernel_degree;       /* GAUSSIAN_PARTICLE_KERNEL_DEGREE (2 for 3DGUT)   */
    float min_kernel_density;    /* GAUSSIAN_PARTICLE_MIN_KERNEL_DENSITY  0.0113    */
    float min_alpha;             /* GAUSSIAN_PARTICLE_MIN_ALPHA           1/255     */
    float max_alpha;             /* GAUSSIAN_PARTICLE_MAX_ALPHA           0.99      */
    float min_transmittance;     /* GAUSSIAN_MIN_TRANSMITTANCE_THRESHOLD  1e-4      */
    float ut_alpha, ut_beta, ut_kappa, ut_delta; /* 1, 2, 0, sqrt(3) */
    float ut_margin;             /* GAUSSIAN_UT_IN_IMAGE_MARGIN_FACTOR    0.1       */
    int32_t rect_bounding, tight_opacity_bounding, tile_culling; /* all 1 */
    int32_t global_z_order;      /* 1 */
    int32_t n_rolling_shutter_iterations; /* GAUSSIAN_N_ROLLING_SHUTTER_ITERATIONS 5 (configs/render/3dgut.yaml:18) */
} gut_oracle_config;

void gut_oracle_default_config(gut_oracle_config* cfg);

/* Host pose maths (sensors.h:44-73): pose at mid exposure, its inverse, view matrix. */
void gut_oracle_sensor_matrices(const gut_oracle_camera* cam,
                                float view_cols[12],   /* world->sensor, 4 columns of 3 (tcnn mat4x3) */
                                float inv_cols[12],    /* sensor->world                                */
                                float cam_pos_world[3]);

uint32_t gut_oracle_higher_msb(uint32_t n);

/* G1: projectOnTiles (gutProjector.cuh:217-322). All outputs are [N,...] float32/uint32. */
void gut_oracle_project(const gut_oracle_config* cfg, const gut_oracle_camera* cam,
                        int64_t n, const float* particles /*[N,12]*/, const float* sph /*[N,48]*/,
                        int32_t sph_degree,
                        uint32_t* tiles_count, float* proj_pos /*[N,2]*/, float* conic_opacity /*[N,4]*/,
                        float* extent /*[N,2]*/, float* depth /*[N]*/, float* rgb /*[N,3]*/,
                        int32_t* visibility /*[N]*/);

/* G2-G5: scan, expand, stable sort on the low (32+higherMsb(T)) bits, tile ranges.
 * keys/values must hold sum(tiles_count) entries; ranges is [T,2]. Returns I. */
int64_t gut_oracle_bin(const gut_oracle_config* cfg, const gut_oracle_camera* cam, int64_t n,
                       const uint32_t* tiles_count, const float* proj_pos, const float* conic_opacity,
                       const float* extent, const float* depth,
                       uint64_t* unsorted_keys, uint32_t* unsorted_values,
                       uint64_t* sorted_keys, uint32_t* sorted_values, uint32_t* ranges);

/* G6: render (gutKBufferRenderer.cuh:274-352, k=0). rays are [H,W,3] in sensor space. */
void gut_oracle_render_forward(const gut_oracle_config* cfg, const gut_oracle_camera* cam,
                               const float* rays_o, const float* rays_d,
                               const float* particles, const float* rgb,
                               const uint32_t* sorted_values, const uint32_t* ranges,
                               float* out_rgba /*[H,W,4]*/, float* out_dist /*[H,W]*/, float* out_hits /*[H,W]*/);

/* G6 with GAUSSIAN_K_BUFFER_SIZE = K > 0 (sorted 3DGUT, gutKBufferRenderer.cuh:62-112,274-352); K <= 64 */
void gut_oracle_render_forward_kbuffer(const gut_oracle_config* cfg, const gut_oracle_camera* cam, int32_t K, const float* rays_o,
                                       const float* rays_d, const float* particles, const float* rgb, const uint32_t* sorted_values,
                                       const uint32_t* ranges, float* out_rgba, float* out_dist, float* out_hits);

/* G7+G8: renderBackward + projectBackward. grads: d_particles [N,12], d_sph [N,48] (zeroed inside). */
void gut_oracle_render_backward(const gut_oracle_config* cfg, const gut_oracle_camera* cam,
                                int64_t n, const float* rays_o, const float* rays_d,
                                const float* particles, const float* sph, int32_t sph_degree,
                                const float* rgb, const uint32_t* tiles_count,
                                const uint32_t* sorted_values, const uint32_t* ranges,
                                const float* out_rgba, const float* out_dist,
                                const float* d_rgba, const float* d_dist,
                                float* d_particles, float* d_sph);

/* adjoint of gut_oracle_render_forward_kbuffer (per-hit adjoint in the buffer's processing order) + G8 */
void gut_oracle_render_backward_kbuffer(const gut_oracle_config* cfg, const gut_oracle_camera* cam, int32_t K, int64_t n,
                                        const float* rays_o, const float* rays_d, const float* particles, const float* sph,
                                        int32_t sph_degree, const float* rgb, const uint32_t* tiles_count,
                                        const uint32_t* sorted_values, const uint32_t* ranges, const float* out_rgba,
                                        const float* out_dist, const float* d_rgba, const float* d_dist, float* d_particles,
                                        float* d_sph);

/* bench-only: render loops visit every k-th tile (bounded CPU sample of a full-size frame); default 1 */
void gut_oracle_set_tile_stride(int k);

/* Single-hit primitives exposed for unit pinning against oracle/_ref. */
int gut_oracle_hit_forward(const gut_oracle_config* cfg, const float ray_o[3], const float ray_d[3],
                           const float particle[12], float* alpha, float* hit_t);
int gut_oracle_hit_backward(const gut_oracle_config* cfg, const float ro[3], const float rd[3], const float p[12],
                            const float prgb[3], float Tint, float* T, float Tgrad, const float Cint[3], float C[3],
                            const float Cgrad[3], float Dint, float* D, float Dgrad, float grad[11], float rgbgrad[3]);
void gut_oracle_sph_eval(int32_t degree, const float coeffs[48], const float dir[3], float rgb_unclamped[3]);

/* ---- 3DGRT (threedgrt_tracer/): brute-force ordered ray tracer, k = 16 hits per trace ----
 * particles/sph as above; rays [R,3] in ray space; ray_to_world = first 3 rows of T_to_world, row major [3,4].
 * out_dist is [R,2] = (integrated distance, last processed hit distance) as in referenceOptix.cu:176-177. */
void grt_oracle_proxies(const gut_oracle_config* cfg, int32_t clamping, int64_t n, const float* particles, float* kscl, float* scene_aabb);
void grt_oracle_trace(const gut_oracle_config* cfg, int32_t clamping, int64_t n, const float* particles, const float* sph,
                      int32_t sph_degree, int64_t n_rays, const float* rays_o, const float* rays_d, const float* ray_to_world,
                      float* out_rgb, float* out_alpha, float* out_dist, float* out_hits, float* visibility);
void grt_oracle_trace_bwd(const gut_oracle_config* cfg, int32_t clamping, int64_t n, const float* particles, const float* sph,
                          int32_t sph_degree, int64_t n_rays, const float* rays_o, const float* rays_d, const float* ray_to_world,
                          const float* out_rgb, const float* out_alpha, const float* out_dist, const float* d_rgb,
                          const float* d_alpha, const float* d_dist, float* d_particles, float* d_sph);

#ifdef __cplusplus
}
#endif
#endif
