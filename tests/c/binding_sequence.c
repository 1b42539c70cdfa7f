/* binding_sequence.c — walks, in plain C against include/rtb.h, exactly the call sequence csharp/RayTracerNative.cs makes
 * through P/Invoke (the C# file cannot be compiled in the build image: no .NET / Unity).  Each block names the C# member it
 * mirrors.  Usage:
 *     binding_sequence host                      no device needed: the host-only calls of the static constructor / marshalling
 *     binding_sequence gpu <scene.txt> <out.rgba> <width> <height>
 *         full sequence on cuda:0; writes the RenderAsync frame of <scene.txt> (row 0 = bottom) for the test to compare.
 * Exit code 0 = every call returned what the C# class expects.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rtb.h"

#define REQUIRE(cond, what)                                                                  \
  do {                                                                                       \
    if (!(cond)) { fprintf(stderr, "FAILED line %d: %s\n", __LINE__, what); return 1; }      \
  } while (0)

static int check(rtb_context* ctx, int rc, const char* what) {
  if (rc != RTB_OK) fprintf(stderr, "%s -> %d: %s\n", what, rc, rtb_last_error(ctx));
  return rc;
}

/* RayTracerNative.ToParams(RenderSettings) */
static rtb_render_params to_params(int w, int h, int depth, int aa) {
  rtb_render_params p;
  rtb_params_default(&p);
  p.has_resolution = 1; p.width = w; p.height = h;
  p.light_intensity = 1.0f;
  p.max_depth = depth;
  p.enable_ambient = p.enable_diffuse = p.enable_specular = p.enable_refraction = 1;
  p.is_orthographic = 0; p.aa_samples = aa;
  return p;
}

/* What EnsureScene marshals from an ObjectData: CSR of transformations, lights, materials, one triangle array with a range per
 * mesh, spheres, boxes — here a hand-made scene of every kind of object. */
typedef struct {
  int32_t offsets[5];
  rtb_xform_elem elems[6];
  int32_t light_xforms[1];
  float light_rgb[3];
  rtb_material materials[2];
  rtb_mesh meshes[1];
  rtb_triangle tris[2];
  rtb_prim spheres[1], boxes[1];
  rtb_scene_desc d;
} HandScene;

static void build_hand_scene(HandScene* s) {
  memset(s, 0, sizeof *s);
  /* T[0] identity; T[1] camera: T(0,0,-74) Rx(-60) Rz(45) (the sample scenes' camera); T[2] light; T[3] sphere / box placement */
  s->offsets[0] = 0; s->offsets[1] = 0;
  s->elems[0] = (rtb_xform_elem){RTB_XF_T, 0, 0, -74, 0};
  s->elems[1] = (rtb_xform_elem){RTB_XF_RX, 0, 0, 0, -60};
  s->elems[2] = (rtb_xform_elem){RTB_XF_RZ, 0, 0, 0, 45};
  s->offsets[2] = 3;
  s->elems[3] = (rtb_xform_elem){RTB_XF_T, -7, 5, 66, 0};
  s->offsets[3] = 4;
  s->elems[4] = (rtb_xform_elem){RTB_XF_T, 0, 0, 4, 0};
  s->elems[5] = (rtb_xform_elem){RTB_XF_S, 6, 6, 6, 0};
  s->offsets[4] = 6;
  s->light_xforms[0] = 2; s->light_rgb[0] = s->light_rgb[1] = s->light_rgb[2] = 1.0f;
  s->materials[0] = (rtb_material){0.8f, 0.3f, 0.2f, 0.1f, 0.7f, 0.3f, 0.0f, 1.0f};
  s->materials[1] = (rtb_material){0.9f, 0.9f, 0.9f, 0.05f, 0.2f, 0.8f, 0.0f, 1.0f};
  s->tris[0] = (rtb_triangle){0, {-20, -20, 0}, {20, -20, 0}, {20, 20, 0}};
  s->tris[1] = (rtb_triangle){0, {-20, -20, 0}, {20, 20, 0}, {-20, 20, 0}};
  s->meshes[0] = (rtb_mesh){0, 0, 0, 2};
  s->spheres[0] = (rtb_prim){3, 1};
  s->boxes[0] = (rtb_prim){3, 0};
  rtb_scene_desc* d = &s->d;
  d->has_image = 1; d->image_w = 200; d->image_h = 200; d->bg[0] = d->bg[1] = d->bg[2] = 0.2f;
  d->has_camera = 1; d->cam_xform = 1; d->cam_distance = 30.0f; d->cam_vfov_deg = 30.0f;
  d->n_xforms = 4; d->xform_offsets = s->offsets; d->xform_elems = s->elems;
  d->n_lights = 1; d->light_xforms = s->light_xforms; d->light_rgb = s->light_rgb;
  d->n_materials = 2; d->materials = s->materials;
  d->n_meshes = 1; d->meshes = s->meshes;
  d->n_triangles = 2; d->triangles = s->tris;
  d->n_spheres = 1; d->spheres = s->spheres;
  d->n_boxes = 1; d->boxes = s->boxes;
}

static int host_only(void) {
  /* static RtbNative(): struct sizes against rtb_abi_sizes */
  int32_t sizes[9] = {0};
  rtb_abi_sizes(sizes, 9);
  const int32_t mine[8] = {(int32_t)sizeof(rtb_xform_elem), (int32_t)sizeof(rtb_material), (int32_t)sizeof(rtb_triangle), (int32_t)sizeof(rtb_mesh),
                           (int32_t)sizeof(rtb_prim), (int32_t)sizeof(rtb_scene_desc), (int32_t)sizeof(rtb_render_params), (int32_t)sizeof(rtb_stats)};
  for (int i = 0; i < 8; i++) REQUIRE(sizes[i] == mine[i], "struct size differs from the library's");
  REQUIRE(rtb_api_version() == RTB_API_VERSION, "api version");
  /* ToParams + Resolve: the library resolves the same resolution the C# Resolve() computes */
  HandScene hs;
  build_hand_scene(&hs);
  rtb_render_params p = to_params(64, 48, 3, 1);
  float uniforms[25];
  int32_t wh[2] = {0, 0};
  REQUIRE(rtb_resolve_frame(&hs.d, &p, uniforms, wh) == RTB_OK && wh[0] == 64 && wh[1] == 48, "rtb_resolve_frame with an override");
  p.has_resolution = 0;
  REQUIRE(rtb_resolve_frame(&hs.d, &p, uniforms, wh) == RTB_OK && wh[0] == 200 && wh[1] == 200, "rtb_resolve_frame from scene.Image");
  hs.d.has_image = 0;
  REQUIRE(rtb_resolve_frame(&hs.d, &p, uniforms, wh) == RTB_OK && wh[0] == 256 && wh[1] == 256, "rtb_resolve_frame default 256 (RayTracer.cs:221-222)");
  return 0;
}

static int gpu_sequence(const char* scene_path, const char* out_path, int w, int h) {
  /* new RayTracerNative() */
  rtb_context* ctx = NULL;
  int rc = rtb_create(&ctx, NULL, 0);
  if (rc != RTB_OK) { fprintf(stderr, "rtb_create -> %d: %s\n", rc, rtb_last_error(NULL)); return 1; }
  int32_t* cancel = (int32_t*)malloc(sizeof(int32_t));  /* Marshal.AllocHGlobal(4) */
  *cancel = 0;

  /* RenderAsync before any scene is known: the C# class always uploads first, the library itself answers RTB_E_NOSCENE */
  rtb_render_params p = to_params(160, 120, 3, 1);
  size_t bytes = (size_t)160 * 120 * 4;
  uint8_t* pixels = (uint8_t*)rtb_alloc_pinned(bytes);
  REQUIRE(pixels != NULL, "rtb_alloc_pinned");
  REQUIRE(rtb_render(ctx, &p, pixels, bytes, NULL, NULL) == RTB_E_NOSCENE, "render before upload must be RTB_E_NOSCENE");

  /* EnsureScene(scene): marshalled ObjectData -> rtb_upload_scene; flags change only after RTB_OK */
  HandScene hs;
  build_hand_scene(&hs);
  REQUIRE(check(ctx, rtb_upload_scene(ctx, &hs.d, RTB_PRIM_TESSELLATED, RTB_BVH_REFERENCE), "rtb_upload_scene") == RTB_OK, "upload");

  /* RenderAsync: cancel flag registered, rtb_render into the Texture2D's raw data, flag cleared */
  int32_t ow = 0, oh = 0;
  REQUIRE(rtb_set_cancel_flag(ctx, cancel) == RTB_OK, "rtb_set_cancel_flag");
  REQUIRE(check(ctx, rtb_render(ctx, &p, pixels, bytes, &ow, &oh), "rtb_render") == RTB_OK && ow == 160 && oh == 120, "RenderAsync");
  REQUIRE(rtb_set_cancel_flag(ctx, NULL) == RTB_OK, "clear cancel flag");
  /* the hand-made scene is in front of the camera: the centre pixel is not background (0.2 -> 51) and alpha is 255 */
  const uint8_t* c = pixels + ((size_t)60 * 160 + 80) * 4;
  REQUIRE(c[3] == 255 && !(c[0] == 51 && c[1] == 51 && c[2] == 51), "centre pixel should show geometry");
  /* GetStats */
  rtb_stats st;
  REQUIRE(rtb_get_stats(ctx, &st) == RTB_OK && st.n_triangles == 2 + 768 + 12 && st.rays_primary == 160 * 120 && st.reserved[0] == 0, "GetStats");
  REQUIRE(st.rays_traversed > 0 && st.rays_traversed <= st.rays_primary + st.rays_continuation + st.rays_shadow, "rays_traversed");

  /* a token cancelled before the call: the C# class returns null without rendering; cancelled DURING the call the library answers -5 */
  *cancel = 1;
  REQUIRE(rtb_set_cancel_flag(ctx, cancel) == RTB_OK, "set flag");
  REQUIRE(rtb_render(ctx, &p, pixels, bytes, NULL, NULL) == RTB_E_CANCELLED, "a raised flag must give RTB_E_CANCELLED");
  *cancel = 0;
  REQUIRE(rtb_set_cancel_flag(ctx, NULL) == RTB_OK, "clear flag");

  /* RenderToTexture, host path: rtb_render_begin / rtb_render_end into the staging texture, every Unity frame */
  uint8_t* staging = (uint8_t*)rtb_alloc_pinned(bytes);
  for (int frame = 0; frame < 3; frame++) {
    int32_t ticket = -1;
    p.has_fov = 1; p.fov_deg = 25.0f + 5.0f * (float)frame;
    REQUIRE(check(ctx, rtb_render_begin(ctx, &p, staging, bytes, &ticket), "rtb_render_begin") == RTB_OK && ticket == frame, "RenderToTexture begin");
    REQUIRE(check(ctx, rtb_render_end(ctx, ticket), "rtb_render_end") == RTB_OK, "RenderToTexture end");
  }
  /* the blocking call with the last settings must give the same frame */
  REQUIRE(rtb_render(ctx, &p, pixels, bytes, NULL, NULL) == RTB_OK && memcmp(pixels, staging, bytes) == 0, "begin/end frame differs from rtb_render");
  p.has_fov = 0;

  /* RenderBegin / RenderEnd with several tickets in flight */
  {
    enum { N = 5 };
    uint8_t* bufs[N];
    int32_t tickets[N];
    for (int k = 0; k < N; k++) { bufs[k] = (uint8_t*)rtb_alloc_pinned(bytes); REQUIRE(bufs[k], "pinned"); }
    for (int k = 0; k < N; k++) REQUIRE(rtb_render_begin(ctx, &p, bufs[k], bytes, &tickets[k]) == RTB_OK, "RenderBegin");
    for (int k = 0; k < N; k++) REQUIRE(rtb_render_end(ctx, tickets[k]) == RTB_OK, "RenderEnd");
    REQUIRE(rtb_render(ctx, &p, pixels, bytes, NULL, NULL) == RTB_OK, "render");
    for (int k = 0; k < N; k++) { REQUIRE(memcmp(bufs[k], pixels, bytes) == 0, "pipelined frame differs"); rtb_free_pinned(bufs[k]); }
  }

  /* InvalidateBVHCache: needsRebuild = true + rtb_invalidate; the next RenderAsync re-uploads the same object */
  REQUIRE(rtb_invalidate(ctx) == RTB_OK, "rtb_invalidate");
  REQUIRE(rtb_render(ctx, &p, pixels, bytes, NULL, NULL) == RTB_E_NOSCENE, "render after invalidate must be RTB_E_NOSCENE");
  REQUIRE(rtb_upload_scene(ctx, &hs.d, RTB_PRIM_TESSELLATED, RTB_BVH_LBVH) == RTB_OK, "re-upload (BvhMode = 1)");
  REQUIRE(rtb_render(ctx, &p, staging, bytes, NULL, NULL) == RTB_OK, "render after re-upload");
  /* ClearRenderTarget / ReleaseBuffers */
  REQUIRE(rtb_clear_target(ctx) == RTB_OK, "rtb_clear_target");
  REQUIRE(rtb_render(ctx, &p, staging, bytes, NULL, NULL) == RTB_OK, "render after ClearRenderTarget");
  REQUIRE(rtb_invalidate(ctx) == RTB_OK && rtb_clear_target(ctx) == RTB_OK, "ReleaseBuffers");

  /* a scene that came from SceneService.LoadScene, rendered like OnStartRayTracingClicked does: the frame the test compares */
  rtb_scene* parsed = NULL;
  char err[256] = "";
  REQUIRE(rtb_scene_load(scene_path, &parsed, err, sizeof err) == RTB_OK, err);
  REQUIRE(check(ctx, rtb_upload_scene(ctx, rtb_scene_get(parsed), RTB_PRIM_TESSELLATED, RTB_BVH_REFERENCE), "rtb_upload_scene(file)") == RTB_OK, "upload");
  rtb_render_params pf = to_params(w, h, 3, 1);
  size_t fbytes = (size_t)w * (size_t)h * 4;
  uint8_t* frame = (uint8_t*)rtb_alloc_pinned(fbytes);
  REQUIRE(frame != NULL, "pinned");
  REQUIRE(check(ctx, rtb_render(ctx, &pf, frame, fbytes, &ow, &oh), "rtb_render(file scene)") == RTB_OK && ow == w && oh == h, "render");
  FILE* f = fopen(out_path, "wb");
  REQUIRE(f && fwrite(frame, 1, fbytes, f) == fbytes && fclose(f) == 0, "write frame");

  /* RenderRotationGif: 36 x 10 degrees in the C# class; 3 frames suffice for the sequence */
  char gif_path[1024];
  snprintf(gif_path, sizeof gif_path, "%s.gif", out_path);
  rtb_render_params pg = to_params(64, 48, 2, 1);
  REQUIRE(check(ctx, rtb_gif_render_rotation(ctx, &pg, 3, 10.0f, gif_path, 10, 0), "rtb_gif_render_rotation") == RTB_OK, "RenderRotationGif");

  /* Dispose */
  rtb_scene_free(parsed);
  rtb_free_pinned(frame); rtb_free_pinned(staging); rtb_free_pinned(pixels);
  rtb_destroy(ctx);
  free(cancel);
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 2 && strcmp(argv[1], "host") == 0) return host_only();
  if (argc == 6 && strcmp(argv[1], "gpu") == 0) {
    if (host_only() != 0) return 1;
    return gpu_sequence(argv[2], argv[3], atoi(argv[4]), atoi(argv[5]));
  }
  fprintf(stderr, "usage: %s host | gpu <scene.txt> <out.rgba> <width> <height>\n", argv[0]);
  return 2;
}
