// host_mirror_test.cpp — exercises include/rtb_raytracer.hpp (the C++ mirror of the reference's RayTracer / ObjectData /
// RenderSettings / SceneService) the way SceneBuilder.cs uses the reference classes (SceneBuilder.cs:489-499, 540-618).
//
//   host_mirror_test host  <scene.txt>                 host-only checks (no GPU): parse, pack, resolve; RayTracer() must fail loudly
//   host_mirror_test render <scene.txt> <out.rgba> W H depth   RenderAsync -> raw RGBA8 (row 0 = bottom), then cache checks
//   host_mirror_test gif    <scene.txt> <out.gif>  W H depth   GifGenerator: 36-frame sweep -> SaveGif, and the fused RenderRotationGif
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>

#include "../../include/rtb_raytracer.hpp"

using namespace rtb;

#define REQUIRE(c)                                                                 \
  do {                                                                             \
    if (!(c)) { std::fprintf(stderr, "REQUIRE failed: %s (line %d)\n", #c, __LINE__); return 1; } \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage\n"); return 2; }
  const std::string mode = argv[1];
  ObjectData scene = SceneService::LoadScene(argv[2]);  // SceneBuilder.cs:492
  REQUIRE(!scene.Transformations.empty() && scene.Camera && scene.Image);
  REQUIRE(SceneService::LoadScene("/nonexistent/scene.txt").TriangleMeshes.empty());  // missing file -> empty ObjectData

  RenderSettings settings;  // defaults = GetRenderSettingsFromUI, SceneBuilder.cs:404-487
  REQUIRE(settings.MaxDepth == 2 && settings.EnableAmbient && settings.AASamples == 1);
  PackedScene packed(scene);
  int32_t wh[2] = {0, 0};
  const rtb_render_params p0 = settings.ToParams();
  REQUIRE(rtb_resolve_frame(packed.desc(), &p0, nullptr, wh) == RTB_OK);
  REQUIRE(wh[0] == scene.Image->horizontal && wh[1] == scene.Image->vertical);  // RayTracer.cs:221-222

  if (mode == "host") {
    try {
      RayTracer rt;
      (void)rt;
      std::printf("a CUDA device is present\n");
    } catch (const Error& e) {
      REQUIRE(e.code == RTB_E_CUDA);  // no CPU fallback
      std::printf("no CUDA device: %s\n", e.what());
    }
    std::printf("host OK %d transformations, %zu meshes\n", (int)scene.Transformations.size(), scene.TriangleMeshes.size());
    return 0;
  }

  REQUIRE(argc >= 7);
  settings.ResolutionOverride = std::array<int, 2>{std::atoi(argv[4]), std::atoi(argv[5])};
  settings.MaxDepth = std::atoi(argv[6]);
  if (mode == "gif") {  // SceneBuilder.OnGifClicked, SceneBuilder.cs:965-1030
    RayTracer rt;
    GifGenerator gen(rt, &scene);
    settings.CameraPositionOverride = Vector3{0, 0, 0};
    settings.CameraRotationOverride = Vector3{-60, 0, 0};
    int reports = 0;
    std::vector<Texture2D> frames = gen.GenerateRotationFrames(settings, [&](float, const std::string&) { reports++; });
    REQUIRE(frames.size() == 36 && reports == 36);
    gen.SaveGif(frames, argv[3]);
    const std::string fused = std::string(argv[3]) + ".fused";
    gen.RenderRotationGif(settings, fused);
    std::ifstream a(argv[3], std::ios::binary), b(fused, std::ios::binary);
    const std::string sa((std::istreambuf_iterator<char>(a)), std::istreambuf_iterator<char>()), sb((std::istreambuf_iterator<char>(b)), std::istreambuf_iterator<char>());
    REQUIRE(sa.size() > 800 && sa == sb);
    gen.SaveGif({}, std::string(argv[3]) + ".none");  // empty list: no file (GifGenerator.cs:162)
    REQUIRE(!std::ifstream(std::string(argv[3]) + ".none").good());
    std::printf("gif OK %zu bytes\n", sa.size());
    return 0;
  }
  RayTracer rt;
  REQUIRE(!rt.RenderAsync(nullptr, settings));  // nothing to render -> null, like the reference
  auto tex = rt.RenderAsync(&scene, settings);
  REQUIRE(tex && tex->width == std::atoi(argv[4]) && tex->height == std::atoi(argv[5]));
  std::ofstream(argv[3], std::ios::binary).write((const char*)tex->pixels.data(), (std::streamsize)tex->pixels.size());
  const float upload_ms = rt.Stats().ms_upload;
  auto again = rt.RenderAsync(&scene, settings);  // BVH cache hit: no re-upload (RayTracer.cs:118-123)
  REQUIRE(again && again->pixels == tex->pixels && rt.Stats().ms_upload == upload_ms);
  rt.InvalidateBVHCache();
  auto third = rt.RenderAsync(&scene, settings);
  REQUIRE(third && third->pixels == tex->pixels);
  volatile int32_t cancel = 1;
  REQUIRE(!rt.RenderAsync(&scene, settings, &cancel));  // cancelled -> null (RayTracer.cs:283)
  auto dev = rt.RenderToTexture(&scene, settings);
  REQUIRE(dev && dev->device_ptr && dev->width == tex->width);
  rt.ReleaseBuffers();
  auto fourth = rt.RenderAsync(&scene, settings);
  REQUIRE(fourth && fourth->pixels == tex->pixels);
  std::printf("render OK %dx%d\n", tex->width, tex->height);
  return 0;
}
