/* rtb.h — C ABI of librtb200.so, the B200-native replacement for the per-pixel render path of
 * mpoboas/cosig-raytracing.
 *
 * The reference has no FFI: its render path is the C# class `RayTracer` (Assets/Services/RayTracer.cs:17) which drives a
 * Unity compute shader (Assets/Shaders/BVHRayTracing.compute).  This header is the plugin interface a maintainer binds with
 * P/Invoke ([DllImport("rtb200")]) in place of ComputeShader.Set / Dispatch / ReadPixels.  Each entry
 * point cites the reference member it replaces.  Plain pointers and sizes only; no exceptions cross the boundary; every
 * function returns an rtb_status (0 = OK).  A context is thread-compatible (one caller at a time, like the reference's
 * main-thread-only RayTracer); calls are blocking unless stated otherwise.
 *
 * All paths in comments are relative to the reference repository root.
 */
#ifndef RTB_H
#define RTB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTB_API_VERSION 1

typedef struct rtb_context rtb_context; /* opaque: owns streams, device memory, peer mappings */

typedef enum rtb_status {
  RTB_OK = 0,
  RTB_E_ARG = -1,       /* null / out-of-range argument */
  RTB_E_CUDA = -2,      /* CUDA runtime failure (no device, OOM, launch error): rtb_last_error has the text */
  RTB_E_NOSCENE = -3,   /* render before upload (reference: null shader -> null, RayTracer.cs:84-88) */
  RTB_E_SIZE = -4,      /* output buffer too small */
  RTB_E_CANCELLED = -5, /* cancel flag observed (reference: token.IsCancellationRequested, RayTracer.cs:283) */
  RTB_E_IO = -6,        /* scene file missing/unreadable */
  RTB_E_PARSE = -7      /* scene file malformed (reference would throw FormatException) */
} rtb_status;

/* TransformType, Assets/Models/ObjectData.cs:66-73 — same enumerator order. */
typedef enum rtb_xform_type { RTB_XF_T = 0, RTB_XF_RX = 1, RTB_XF_RY = 2, RTB_XF_RZ = 3, RTB_XF_S = 4 } rtb_xform_type;

/* TransformElement, ObjectData.cs:80-90.  x,y,z used by T and S; angle_deg by Rx/Ry/Rz. */
typedef struct rtb_xform_elem { int32_t type; float x, y, z; float angle_deg; } rtb_xform_elem;

/* MaterialDescription, ObjectData.cs:157-176 == GPUMaterial (32 B), RayTracer.cs:442-450. */
typedef struct rtb_material { float r, g, b; float ka, kd, ks, kr, ior; } rtb_material;

/* Triangle, ObjectData.cs:196-215. */
typedef struct rtb_triangle { int32_t material; float v0[3], v1[3], v2[3]; } rtb_triangle;

/* TrianglesMesh, ObjectData.cs:183-190: a range of `triangles` sharing one transformation. */
typedef struct rtb_mesh { int32_t xform; int32_t reserved; int64_t first_tri; int64_t n_tris; } rtb_mesh;

/* SphereDescription / BoxDescription, ObjectData.cs:221-241. */
typedef struct rtb_prim { int32_t xform; int32_t material; } rtb_prim;

/* ObjectData, ObjectData.cs:9-34.  The library copies everything it needs during rtb_upload_scene; the caller keeps
 * ownership of all pointed-to memory. */
typedef struct rtb_scene_desc {
  int32_t has_image; int32_t image_w, image_h; float bg[3];                /* ImageSettings  :40-50  */
  int32_t has_camera; int32_t cam_xform; float cam_distance, cam_vfov_deg; /* CameraSettings :128-138 */
  int32_t n_xforms;                    /* CompositeTransformation list as CSR: elements of transform i are   */
  const int32_t* xform_offsets;        /*   xform_elems[xform_offsets[i] .. xform_offsets[i+1])  (n_xforms+1) */
  const rtb_xform_elem* xform_elems;
  int32_t n_lights; const int32_t* light_xforms; const float* light_rgb;   /* LightSource :144-151 (rgb: 3 per light) */
  int32_t n_materials; const rtb_material* materials;
  int32_t n_meshes; const rtb_mesh* meshes;
  int64_t n_triangles; const rtb_triangle* triangles;
  int32_t n_spheres; const rtb_prim* spheres;
  int32_t n_boxes; const rtb_prim* boxes;
} rtb_scene_desc;

/* How geometry is intersected. */
enum { RTB_PRIM_TESSELLATED = 0, /* reference behaviour: spheres -> 768 tris, boxes -> 12 tris (SceneGeometryConverter.cs) */
       RTB_PRIM_ANALYTIC = 1 };  /* unit-sphere quadratic / unit-box slabs (semantics of Services/BVH/HittableObjects.cs)   */
/* Which acceleration structure is traversed. */
enum { RTB_BVH_REFERENCE = 0,    /* host median-split build identical to BVHBuilder.cs, reference traversal order: ids bit-exact */
       RTB_BVH_LBVH = 1 };       /* Morton-code LBVH built on the GPU, ordered traversal (tie winners may differ)               */
/* Output addressing for the device-pointer entry point. */
enum { RTB_OUT_FRAME = 0,        /* dst is a full W*H frame; only this rank's rows are written (peer-store gather)    */
       RTB_OUT_COMPACT = 1 };    /* dst holds this rank's bands packed in local row order (NCCL gather staging)       */

/* RenderSettings, Assets/Models/RenderSettings.cs:7-70 (nullable -> has_*), plus the knobs this library adds. */
typedef struct rtb_render_params {
  int32_t has_resolution, width, height;          /* ResolutionOverride       */
  int32_t has_bg; float bg[3];                    /* BackgroundColorOverride  */
  float light_intensity;                          /* LightIntensityScale      */
  int32_t has_cam_pos; float cam_pos[3];          /* CameraPositionOverride   */
  int32_t has_cam_rot; float cam_rot_euler_deg[3];/* CameraRotationOverride   */
  int32_t has_fov; float fov_deg;                 /* CameraFovOverride        */
  int32_t max_depth;                              /* MaxDepth                 */
  int32_t enable_ambient, enable_diffuse, enable_specular, enable_refraction;
  int32_t is_orthographic;
  int32_t aa_samples;
  int32_t soft_shadows; float light_size;
  int32_t glossy; float roughness;
  int32_t motion_blur; float shutter_speed;
  /* --- additions (all zero = reference behaviour) --- */
  int32_t debug_mode;        /* _DebugMode 0..3, BVHRayTracing.compute:484-508; the reference host always passes 0 */
  int32_t srgb_encode;       /* 0: byte = floor(saturate(c)*255+0.5) */
  int32_t band_rank, band_world, band_rows; /* process-per-GPU tile sharding: bands of band_rows rows, band b -> rank b % world.
                                               band_world <= 1 means the whole frame.  band_rows 0 -> 32.               */
  int32_t out_layout;        /* RTB_OUT_*; only used by rtb_render_device */
  int32_t reserved[6];
} rtb_render_params;

/* Counters of the last render call on this context (rays are TraverseBVH-equivalent queries). */
typedef struct rtb_stats {
  int64_t rays_primary, rays_continuation, rays_shadow;
  int64_t paths_hit_primary;
  int64_t n_triangles, n_nodes;
  int32_t width, height, spp, chunks;
  int32_t kernel_launches;        /* launches of this library's kernels in the last render */
  int32_t n_devices;
  float ms_upload, ms_build;      /* last rtb_upload_scene: host prep + H2D, device build (flatten + BVH) */
  float ms_render_device;         /* last render: first launch -> last kernel, CUDA events, max over devices */
  float ms_traverse, ms_shade, ms_resolve; /* per kernel family (k_traverse = all BVH queries, k_shade, k_resolve), summed over
                                             depths/chunks on device 0; only when profiling is enabled */
  int64_t h2d_bytes, d2h_bytes;   /* bytes copied across PCIe by the last render call */
  int64_t reserved[4];            /* [0] traversal-stack overflows (must be 0) [1] BVH node visits (per ray) [2] triangles tested (per ray)
                                     [3] diagnostic builds: node visits of the longest ray */
  int64_t rays_traversed;         /* rays that entered the BVH: all rays minus the primary rays k_raygen resolved against the scene's
                                     root box (they count as rays — the reference traces them — but cost one box test) */
  int64_t packet_node_fetches, packet_tri_fetches; /* record fetches of the packet kernels (one per warp and visit), when enabled */
  int64_t bytes_per_slot;         /* wavefront queue state per pixel-sample slot and lane */
} rtb_stats;

/* new RayTracer() + SetComputeShader, RayTracer.cs:17-32.  device_ids == NULL -> {0}.  With n_devices > 1 the frame is
 * sharded by row bands over the devices of this process and gathered on device_ids[0] by peer stores over NVLink. */
int rtb_create(rtb_context** out, const int32_t* device_ids, int32_t n_devices);

/* ReleaseBuffers + destruction, RayTracer.cs:47-59. */
void rtb_destroy(rtb_context* ctx);

/* Fills `p` with the reference UI defaults (SceneBuilder.cs:335-343,401,439-445): depth 2, all lighting toggles on,
 * intensity 1, AA 1, no overrides, tessellated primitives, reference BVH. */
void rtb_params_default(rtb_render_params* p);

/* RebuildBVH, RayTracer.cs:386-404 (ExtractTriangles + BVHBuilder.Build + SetData) and SetupMaterialBuffer :455-499.
 * primitive_mode / bvh_mode are RTB_PRIM_* / RTB_BVH_*.  Geometry is flattened and (for LBVH) the hierarchy built on
 * the device(s).  An empty scene is legal and renders the background. */
int rtb_upload_scene(rtb_context* ctx, const rtb_scene_desc* scene, int32_t primitive_mode, int32_t bvh_mode);

/* InvalidateBVHCache, RayTracer.cs:38-42: drops device geometry; the next render returns RTB_E_NOSCENE until re-upload. */
int rtb_invalidate(rtb_context* ctx);

/* ClearRenderTarget, RayTracer.cs:65-72: frees the cached frame/queue buffers (they are re-created on demand). */
int rtb_clear_target(rtb_context* ctx);

/* RenderAsync, RayTracer.cs:212-380 (blocking).  Writes width*height RGBA8 pixels, row 0 = bottom of the picture
 * (Unity Texture2D convention), into caller memory (host; pinned memory from rtb_alloc_pinned avoids
 * a staging copy).  `bytes` is the capacity of rgba8.  out_w/out_h (optional) receive the resolved resolution. */
int rtb_render(rtb_context* ctx, const rtb_render_params* p, uint8_t* rgba8, size_t bytes, int32_t* out_w, int32_t* out_h);

/* Pipelined RenderAsync for hosts that render a stream of frames (the reference's realtime mode calls its renderer once per
 * Unity frame, SceneBuilder.cs:521-537): rtb_render_begin enqueues the frame and its readback into `rgba8` (page-locked memory,
 * see rtb_alloc_pinned, for a truly asynchronous copy) and returns a ticket; rtb_render_end blocks until that frame is in
 * `rgba8`.  Up to 16 frames may be in flight; `rgba8` must stay valid until its rtb_render_end.  Multi-device contexts
 * (rtb_create with several device ids) pipeline the same way: every device renders its bands of frame k+1 while device 0
 * copies frame k out. */
int rtb_render_begin(rtb_context* ctx, const rtb_render_params* p, uint8_t* rgba8, size_t bytes, int32_t* ticket);
int rtb_render_end(rtb_context* ctx, int32_t ticket);

/* RenderToTexture, RayTracer.cs:82-202 (no readback): renders into device memory `dst_device` (on device_ids[0], or a
 * peer-mapped pointer into another GPU's frame for the fused NVLink gather).  Asynchronous on the context's stream unless
 * `sync` != 0.  Honors band_rank/band_world/out_layout. */
int rtb_render_device(rtb_context* ctx, const rtb_render_params* p, void* dst_device, size_t bytes, int32_t sync);

/* Primary-hit maps for parity checks (the reference exposes them only as debug views, BVHRayTracing.compute:484-508):
 * per pixel (row 0 = bottom) the closest hit of the pixel-centre ray: prim_id = triangle index in ExtractTriangles
 * emission order (-1 = miss), t, material index.  Any pointer may be NULL.  Host pointers. */
int rtb_render_aux(rtb_context* ctx, const rtb_render_params* p, int32_t* prim_id, float* t, int32_t* material);

/* Copies the flattened object-space triangle arrays back to the host for parity checks against the oracle:
 * for triangle i in emission order: v0,v1,v2,n0,n1,n2 (18 floats) and material.  Either pointer may be NULL;
 * n_out receives the triangle count. */
int rtb_get_triangles(rtb_context* ctx, float* v_n_18, int32_t* material, int64_t capacity, int64_t* n_out);

int rtb_get_stats(rtb_context* ctx, rtb_stats* out);
int rtb_set_profiling(rtb_context* ctx, int32_t enable); /* per-kernel-family CUDA-event timing in rtb_stats */
int rtb_set_cancel_flag(rtb_context* ctx, const volatile int32_t* flag); /* polled between wavefront depths */
int rtb_synchronize(rtb_context* ctx);
/* Asynchronous renders (rtb_render_device with sync == 0) rotate over a few internal streams per device (RTB_LANES, default 6)
 * so that successive chunks / frames overlap.  rtb_get_stream returns the primary stream (cudaStream_t) of device `index`;
 * rtb_flush makes that stream wait for everything enqueued so far on the others, so that an event the host records on it
 * afterwards (or work it orders behind it) covers all frames in flight. */
void* rtb_get_stream(rtb_context* ctx, int32_t index);
int rtb_flush(rtb_context* ctx);

/* Context-owned UTF-8 text for the last failing call on this context (ctx may be NULL for rtb_create failures). */
const char* rtb_last_error(rtb_context* ctx);

/* Page-locked host memory for rgba8 outputs (C# side: wrap with NativeArray / LoadRawTextureData). */
void* rtb_alloc_pinned(size_t bytes);
void rtb_free_pinned(void* p);

/* CUDA IPC plumbing for process-per-GPU hosts (bench.py under torchrun): export the context's device-0 frame buffer of
 * `bytes` bytes (allocated on demand) as a 64-byte handle; open a peer's handle and get a device pointer usable as
 * `dst_device`.  Opened pointers are closed by rtb_destroy.  handle64 == NULL: no export, the call only returns the pointer of the
 * context's own frame buffer.  Once exported the buffer neither moves nor shrinks: a later call that would need a larger one, and
 * rtb_clear_target, fail with RTB_E_ARG instead of leaving importers with a dangling mapping. */
int rtb_frame_export(rtb_context* ctx, size_t bytes, void** dev_ptr, uint8_t handle64[64]);
int rtb_frame_import(rtb_context* ctx, const uint8_t handle64[64], void** dev_ptr);
/* Process-per-GPU rendering of ONE frame stream by `world` ranks (one process and one single-device context each): the
 * frame is split into bands of params.band_rows rows (default 8), band b is rendered by rank b % world, every rank's resolve kernel
 * stores its bands straight into a frame buffer of rank 0 over NVLink, and rank 0 reads whole frames back — pipelined over
 * `n_buffers` (1..8) frame buffers with device-side sequence flags, so no rank ever blocks on another from the host.
 * rtb_group_create: rank 0 allocates the ring (n_buffers frames of frame_bytes + flags) and FILLS handle64; every other rank
 * receives those 64 bytes from the host's own channel (MPI, torch.distributed, a pipe) and PASSES them in.  All ranks then call
 * rtb_group_render_begin for the same sequence of frames (band_rank / band_world / out_layout of `p` are set by the library);
 * `rgba8` (pinned host memory, width*height*4 bytes) is used on rank 0 only.  rtb_group_render_end: rank 0 — the frame is in
 * its host buffer; other ranks — their bands are stored.  Up to 16 frames may be begun before the oldest is ended.  A rank that
 * does not arrive within RTB_GROUP_TIMEOUT_MS (default 10000) turns into RTB_E_CUDA on the waiting ranks, never a hang.
 * Destroy the group (or the contexts) on every rank before rank 0's context goes away. */
int rtb_group_create(rtb_context* ctx, int32_t rank, int32_t world, size_t frame_bytes, int32_t n_buffers, uint8_t handle64[64]);
int rtb_group_render_begin(rtb_context* ctx, const rtb_render_params* p, uint8_t* rgba8, size_t bytes, int32_t* ticket);
int rtb_group_render_end(rtb_context* ctx, int32_t ticket);
int rtb_group_destroy(rtb_context* ctx);
/* Host-ring form of the same group, for consumers of the frames in HOST memory: the ring of `n_buffers` frames lies in POSIX shared
 * memory `shm_name` ("/name": rank 0 creates it, the other ranks attach after rank 0's call has returned; the name travels over
 * the host's own channel), page-locked by every rank.  Every rank renders its bands into a device buffer of its own and copies
 * exactly those bands into the frame's slot over ITS OWN PCIe link (one strided copy), so the readback bandwidth grows with the
 * number of GPUs instead of being rank 0's link alone; no device memory is shared between the processes.  Same begin / end calls,
 * with rgba8 = NULL on every rank; after rtb_group_render_end(ticket) rank 0 finds the whole frame at rtb_group_frame(ticket),
 * where it stays until rank 0 begins frame ticket + n_buffers (the other ranks do not overwrite a slot before that: they wait,
 * on the host, for rank 0's begin of the same frame, bounded by RTB_GROUP_TIMEOUT_MS).  Rank 0 therefore keeps at most n_buffers
 * frames begun and not ended: beginning one more is RTB_E_ARG (it would overwrite a frame nobody has seen). */
int rtb_group_create_host(rtb_context* ctx, int32_t rank, int32_t world, size_t frame_bytes, int32_t n_buffers, const char* shm_name);
int rtb_group_frame(rtb_context* ctx, int32_t ticket, const uint8_t** rgba8);

/* ReadPixels (RayTracer.cs:371-375) of the context's own frame buffer — the one rtb_frame_export shares — after the
 * peers have stored their bands into it. */
int rtb_frame_read(rtb_context* ctx, uint8_t* rgba8, size_t bytes);

/* Zero-copy display: the reference's realtime mode never reads the frame back — RenderToTexture returns the GPU
 * RenderTexture (RayTracer.cs:82-202) and UI Toolkit shows it (SceneBuilder.cs:836-852).  The counterpart here: the host's graphics
 * API allocates the buffer its texture is fed from, exports it (Vulkan vkGetMemoryFdKHR / vkGetMemoryWin32HandleKHR, D3D12
 * CreateSharedHandle), and rtb_external_import maps it into device_ids[0]'s address space (cudaImportExternalMemory); the
 * returned pointer is then a valid `dst_device` of rtb_render_device, whose resolve kernel stores the RGBA8 pixels straight into
 * it.  `handle`: the file descriptor cast to a pointer (RTB_EXT_OPAQUE_FD; owned by CUDA after a successful import — do not
 * close it) or the NT handle.  `dedicated` != 0 for D3D12 committed resources / Vulkan dedicated allocations.  Synchronise with
 * the graphics queue by `sync` = 1 (or rtb_synchronize) before the graphics API samples the buffer.  rtb_external_release (or
 * rtb_destroy) unmaps. */
enum { RTB_EXT_OPAQUE_FD = 1, RTB_EXT_OPAQUE_WIN32 = 2, RTB_EXT_D3D12_HEAP = 4, RTB_EXT_D3D12_RESOURCE = 5 };
int rtb_external_import(rtb_context* ctx, int32_t handle_type, void* handle, size_t bytes, int32_t dedicated, void** dev_ptr);
int rtb_external_release(rtb_context* ctx, void* dev_ptr);

/* Parity/debug access to the acceleration structure of the uploaded scene: nodes as stored on the device (reference
 * mode: 8 words per node = GPUBVHNode, BVHBuilder.cs:27-34; LBVH mode: 16 words per node) and the
 * leaf-order -> emission-order triangle permutation (the reference declares it as BVHResult.triangleIndices,
 * BVHBuilder.cs:67, but never fills it). */
int rtb_get_bvh(rtb_context* ctx, void* nodes, int64_t nodes_capacity_bytes, int64_t* n_nodes, int32_t* perm, int64_t perm_capacity);
/* 32-bit words per node record of the uploaded scene as rtb_get_bvh returns them: 8 (reference mode), 24 (LBVH mode: 8-wide
 * quantised records) or 16 (LBVH mode with RTB_WIDE=0: binary two-box records); 0 without a scene. */
int rtb_get_bvh_node_words(rtb_context* ctx);

/* Host-only helpers (no CUDA device needed; used by the CPU test-suite and by hosts that want the uniforms).
 * rtb_resolve_frame: what RayTracer.cs:221-355 resolves from (scene, settings): out25 = cameraToObject (row-major 4x4),
 *   _CameraDistance, tan(fov/2), _OrthoSize, _LightPosition xyz, _BackgroundColor rgb; wh = width, height.
 * rtb_build_reference_bvh: BVHBuilder.Build (BVHBuilder.cs:76-95) on caller-provided triangles, 12 floats each
 *   (v0.xyz, c.x, v1.xyz, c.y, v2.xyz, c.z); nodes8 = 8 words per node, perm = leaf order -> input index.
 * rtb_abi_sizes: sizeof of the 8 public structs in declaration order, for binding self-checks; entry [8] = 32-bit words per
 *   LBVH node record as returned by rtb_get_bvh. */
int rtb_resolve_frame(const rtb_scene_desc* scene, const rtb_render_params* p, float* out25, int32_t* wh);
int rtb_build_reference_bvh(const float* raw12, int32_t n, float* nodes8, int64_t nodes_capacity, int64_t* n_nodes, int32_t* perm);
void rtb_abi_sizes(int32_t* out, int32_t n);

/* SceneService.LoadScene, Assets/Services/SceneService.cs:26-242: parses the COSIG scene text format.  The returned
 * object owns its arrays; rtb_scene_get gives a desc view valid until rtb_scene_free. */
typedef struct rtb_scene rtb_scene;
int rtb_scene_load(const char* path, rtb_scene** out, char* err, size_t err_cap);
int rtb_scene_parse(const char* text, size_t len, rtb_scene** out, char* err, size_t err_cap);
const rtb_scene_desc* rtb_scene_get(const rtb_scene* s);
void rtb_scene_free(rtb_scene* s);

/* ---- GIF sweep: GifGenerator, Assets/Services/GifGenerator.cs -------------------------------------------------
 * The reference renders 36 frames (camera rotation override Z = 0, 10, ... 350 degrees, :49-63), maps every pixel to a
 * 6x6x6 colour-cube index with a vertical flip (ConvertToIndexed :346-369), LZW-compresses each frame (:411-501, Parallel.For
 * over frames :123-130) and writes GIF89a (:82-155).  Here the palette mapping is a device kernel, so a frame leaves the GPU
 * as 1 byte per pixel; LZW runs on host threads while later frames render.  Output files are byte-identical to the reference
 * algorithm's (tests/test_gif_*.py). */
void rtb_gif_color_table(uint8_t rgb768[768]);                       /* GenerateColorTable :219-247 */
int64_t rtb_gif_lzw_bound(int64_t n);                                /* capacity rtb_gif_lzw needs for n input bytes */
int64_t rtb_gif_lzw(const uint8_t* indexed, int64_t n, uint8_t* out, int64_t capacity); /* LzwCompress :411-501 (host); bytes written or < 0 */
/* ConvertToIndexed :346-369 on the device: host RGBA8 frame (row 0 = bottom, Texture2D order) -> width*height palette indices,
 * top row first. */
int rtb_gif_index_frame(rtb_context* ctx, const uint8_t* rgba8, int32_t width, int32_t height, uint8_t* indexed);
/* The same for a frame already on device 0 (e.g. the target of rtb_render_device): device pointers, asynchronous on
 * rtb_get_stream(ctx, 0). */
int rtb_gif_index_device(rtb_context* ctx, const void* rgba8_device, int32_t width, int32_t height, void* indexed_device);
/* Like rtb_render_begin, but the frame is converted on the device (ConvertToIndexed) and only the width*height palette
 * indices are read back into `indexed` (top row first).  Shares the ticket ring with rtb_render_begin; wait with rtb_render_end. */
int rtb_render_begin_indexed(rtb_context* ctx, const rtb_render_params* p, uint8_t* indexed, size_t bytes, int32_t* ticket);
/* SaveGifAsync :82-155 for frames given as palette indices (top row first; host-only: LZW + container) / as RGBA8 Texture2D data
 * (row 0 = bottom; the palette mapping runs on the device, so a context is required).  threads <= 0: all host cores. */
int rtb_gif_save_indexed(const char* path, int32_t width, int32_t height, const uint8_t* const* frames, int32_t n_frames,
                         int32_t frame_delay_cs, int32_t threads);
int rtb_gif_save(rtb_context* ctx, const char* path, int32_t width, int32_t height, const uint8_t* const* rgba8_frames, int32_t n_frames,
                 int32_t frame_delay_cs, int32_t threads);
/* GenerateRotationFrames :40-72 + SaveGifAsync fused: frame k renders with CameraRotationOverride = (base.x, base.y, k*step_deg)
 * (the reference: n_frames 36, step 10), is palette-mapped on the device, read back as indices and compressed on a host thread
 * while the following frames render; the file is written when all frames are in.  Honors the cancel flag (RTB_E_CANCELLED). */
int rtb_gif_render_rotation(rtb_context* ctx, const rtb_render_params* base, int32_t n_frames, float step_deg, const char* path,
                            int32_t frame_delay_cs, int32_t threads);

int rtb_api_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RTB_H */
