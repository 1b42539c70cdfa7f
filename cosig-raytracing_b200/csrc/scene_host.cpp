// scene_host.cpp — see scene_host.hpp.  Host-only code (no device work); FP32, one rounding per operation
// (-ffp-contract=off), transcendental inputs through double libm rounded once.
#include "scene_host.hpp"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <stdexcept>

namespace rtb {

// ---------------------------------------------------------------------------------------------------------------------
// HostScene
// ---------------------------------------------------------------------------------------------------------------------
void HostScene::relink() {
  d.n_xforms = (int32_t)xform_offsets.size() - 1;
  d.xform_offsets = xform_offsets.data();
  d.xform_elems = xform_elems.data();
  d.n_lights = (int32_t)light_xforms.size();
  d.light_xforms = light_xforms.data();
  d.light_rgb = light_rgb.data();
  d.n_materials = (int32_t)materials.size();
  d.materials = materials.data();
  d.n_meshes = (int32_t)meshes.size();
  d.meshes = meshes.data();
  d.triangles = triangles.data();
  d.n_spheres = (int32_t)spheres.size();
  d.spheres = spheres.data();
  d.n_boxes = (int32_t)boxes.size();
  d.boxes = boxes.data();
}

std::string HostScene::assign(const rtb_scene_desc& s, bool copy_triangles) {
  if (s.n_xforms < 0 || s.n_lights < 0 || s.n_materials < 0 || s.n_meshes < 0 || s.n_triangles < 0 || s.n_spheres < 0 || s.n_boxes < 0)
    return "negative count in scene description";
  if (s.n_xforms > 0 && (!s.xform_offsets || s.xform_offsets[0] != 0)) return "xform_offsets missing or not starting at 0";
  for (int i = 0; i < s.n_xforms; i++)
    if (s.xform_offsets[i + 1] < s.xform_offsets[i]) return "xform_offsets not monotone";
  const int n_el = s.n_xforms > 0 ? s.xform_offsets[s.n_xforms] : 0;
  if (n_el > 0 && !s.xform_elems) return "xform_elems is null";
  if (s.n_lights > 0 && !s.light_xforms) return "light_xforms is null";
  if (s.n_materials > 0 && !s.materials) return "materials is null";
  if (s.n_meshes > 0 && !s.meshes) return "meshes is null";
  if (s.n_triangles > 0 && !s.triangles) return "triangles is null";
  if (s.n_spheres > 0 && !s.spheres) return "spheres is null";
  if (s.n_boxes > 0 && !s.boxes) return "boxes is null";
  for (int i = 0; i < s.n_meshes; i++) {
    const rtb_mesh& m = s.meshes[i];
    if (m.first_tri < 0 || m.n_tris < 0 || m.first_tri + m.n_tris > s.n_triangles) return "mesh triangle range out of bounds";
  }
  d = s;
  if (s.n_xforms > 0) xform_offsets.assign(s.xform_offsets, s.xform_offsets + s.n_xforms + 1);
  else xform_offsets.assign(1, 0);
  xform_elems.assign(s.xform_elems, s.xform_elems + n_el);
  light_xforms.assign(s.light_xforms, s.light_xforms + s.n_lights);
  if (s.light_rgb) light_rgb.assign(s.light_rgb, s.light_rgb + 3 * (size_t)s.n_lights);
  else light_rgb.assign(3 * (size_t)s.n_lights, 1.0f);
  materials.assign(s.materials, s.materials + s.n_materials);
  meshes.assign(s.meshes, s.meshes + s.n_meshes);
  if (copy_triangles) triangles.assign(s.triangles, s.triangles + s.n_triangles);
  else triangles.clear();
  spheres.assign(s.spheres, s.spheres + s.n_spheres);
  boxes.assign(s.boxes, s.boxes + s.n_boxes);
  relink();
  if (!copy_triangles) d.triangles = nullptr;
  return std::string();
}

// ---------------------------------------------------------------------------------------------------------------------
// Composite transforms and the object table
// ---------------------------------------------------------------------------------------------------------------------
Mat4 composite_matrix(const rtb_scene_desc& s, int index) {
  Mat4 M = Mat4::identity();
  if (index < 0 || index >= s.n_xforms) return M;  // SceneGeometryConverter.cs:85
  for (int k = s.xform_offsets[index]; k < s.xform_offsets[index + 1]; k++) {
    const rtb_xform_elem& e = s.xform_elems[k];
    Mat4 E = Mat4::identity();
    switch (e.type) {
      case RTB_XF_T: E = translate(Vec3f{e.x, e.y, e.z}); break;
      case RTB_XF_S: E = scale(Vec3f{e.x, e.y, e.z}); break;
      case RTB_XF_RX: E = rotate(angle_axis(e.angle_deg, Vec3f{1, 0, 0})); break;
      case RTB_XF_RY: E = rotate(angle_axis(e.angle_deg, Vec3f{0, 1, 0})); break;
      case RTB_XF_RZ: E = rotate(angle_axis(e.angle_deg, Vec3f{0, 0, 1})); break;
      default: break;
    }
    M = M * E;  // appended on the right: the last element listed acts on the point first
  }
  return M;
}

static void fill_object(FlattenObject& o, const Mat4& M, bool with_normal_matrix) {
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 4; c++) o.m[r * 4 + c] = M.at(r, c);
  std::memset(o.nm, 0, sizeof o.nm);
  if (with_normal_matrix) {
    const Mat4 N = transpose(inverse(M));
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) o.nm[r * 3 + c] = N.at(r, c);
  }
}

int64_t build_object_table(const rtb_scene_desc& s, bool analytic, std::vector<FlattenObject>& out, std::vector<float>& prims) {
  out.clear();
  prims.clear();
  int64_t at = 0;
  auto push_analytic = [&](int kind, int xform, int material) {
    const Mat4 M = composite_matrix(s, xform);
    const Mat4 W = inverse(M);
    FlattenObject o;
    std::memset(&o, 0, sizeof o);
    fill_object(o, M, false);
    o.kind = kind;
    o.material = material;
    o.out_first = (int32_t)at;
    o.src_first = (int32_t)(prims.size() / 24);
    o.count = 1;
    for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) prims.push_back(M.at(r, c));
    for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) prims.push_back(W.at(r, c));
    at += 1;
    out.push_back(o);
  };
  auto push = [&](int kind, int xform, int material, int64_t src_first, int64_t count) {
    FlattenObject o;
    std::memset(&o, 0, sizeof o);
    fill_object(o, composite_matrix(s, xform), kind == OBJ_SPHERE);
    o.kind = kind;
    o.material = material;
    o.out_first = (int32_t)at;
    o.src_first = (int32_t)src_first;
    o.count = (int32_t)count;
    at += count;
    if (count > 0) out.push_back(o);
  };
  for (int i = 0; i < s.n_meshes && at <= INT32_MAX; i++) push(OBJ_MESH, s.meshes[i].xform, 0, s.meshes[i].first_tri, s.meshes[i].n_tris);
  for (int i = 0; i < s.n_boxes && at <= INT32_MAX; i++) {
    if (analytic) push_analytic(OBJ_BOX_ANALYTIC, s.boxes[i].xform, s.boxes[i].material);
    else push(OBJ_BOX, s.boxes[i].xform, s.boxes[i].material, 0, kBoxTris);
  }
  for (int i = 0; i < s.n_spheres && at <= INT32_MAX; i++) {
    if (analytic) push_analytic(OBJ_SPHERE_ANALYTIC, s.spheres[i].xform, s.spheres[i].material);
    else push(OBJ_SPHERE, s.spheres[i].xform, s.spheres[i].material, 0, kSphereTris);
  }
  if (at > INT32_MAX - 64 || s.n_triangles > INT32_MAX - 64) return -1;
  return at;
}

const float* unit_sphere_table() {
  static float table[kSphereVerts * 3];
  static bool ready = false;
  if (!ready) {
    const int n_long = 24, n_lat = 16;
    const float pi = 3.14159274f;  // Mathf.PI
    const float two_pi = pi * 2.0f;
    auto put = [&](int i, float x, float y, float z) { table[i * 3] = x; table[i * 3 + 1] = y; table[i * 3 + 2] = z; };
    put(0, 0.0f, 1.0f, 0.0f);
    for (int lat = 0; lat < n_lat; lat++) {
      const float a1 = pi * (float)(lat + 1) / (float)(n_lat + 1);
      const float s1 = (float)std::sin((double)a1), c1 = (float)std::cos((double)a1);
      for (int lon = 0; lon <= n_long; lon++) {
        const float a2 = two_pi * (float)(lon == n_long ? 0 : lon) / (float)n_long;
        const float s2 = (float)std::sin((double)a2), c2 = (float)std::cos((double)a2);
        put(lon + lat * (n_long + 1) + 1, (s1 * c2) * 1.0f, c1 * 1.0f, (s1 * s2) * 1.0f);
      }
    }
    put(kSphereVerts - 1, 0.0f, -1.0f, 0.0f);
    ready = true;
  }
  return table;
}

void pack_materials(const rtb_scene_desc& s, std::vector<float>& out8) {
  out8.clear();
  if (s.n_materials == 0) {  // RayTracer.cs:457-474
    const float def[8] = {1.0f, 1.0f, 1.0f, 0.1f, 0.7f, 0.0f, 0.0f, 1.0f};
    out8.assign(def, def + 8);
    return;
  }
  out8.resize((size_t)s.n_materials * 8);
  for (int i = 0; i < s.n_materials; i++) {
    const rtb_material& m = s.materials[i];
    float* o = &out8[(size_t)i * 8];
    o[0] = m.r; o[1] = m.g; o[2] = m.b; o[3] = m.ka; o[4] = m.kd; o[5] = m.ks; o[6] = m.kr; o[7] = m.ior;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Per-frame uniforms
// ---------------------------------------------------------------------------------------------------------------------
bool resolve_frame(const rtb_scene_desc& s, const rtb_render_params& p, FrameParams& f, std::string& err) {
  std::memset(&f, 0, sizeof f);
  // RayTracer.cs:221-222: override, else scene image (at least 1), else 256
  f.width = p.has_resolution ? p.width : std::max(1, s.has_image ? s.image_w : 256);
  f.height = p.has_resolution ? p.height : std::max(1, s.has_image ? s.image_h : 256);
  if (f.width < 1 || f.height < 1 || f.width > 65536 || f.height > 65536) { err = "resolution out of range"; return false; }
  if (p.max_depth < 0 || p.max_depth > 64) { err = "max_depth must be in [0, 64]"; return false; }
  if (p.aa_samples > 1024) { err = "aa_samples must be <= 1024"; return false; }

  Mat4 scene_cam = Mat4::identity();  // :238-243
  if (s.has_camera && s.cam_xform >= 0 && s.cam_xform < s.n_xforms) scene_cam = composite_matrix(s, s.cam_xform);
  Mat4 cam_to_object;
  if (p.has_cam_pos || p.has_cam_rot) {  // :251-261 — Matrix4x4.TRS(pos, Quaternion.Euler(rot), one).inverse
    const Vec3f pos = p.has_cam_pos ? Vec3f{p.cam_pos[0], p.cam_pos[1], p.cam_pos[2]} : Vec3f{0, 0, 0};
    const Vec3f rot = p.has_cam_rot ? Vec3f{p.cam_rot_euler_deg[0], p.cam_rot_euler_deg[1], p.cam_rot_euler_deg[2]} : Vec3f{0, 0, 0};
    cam_to_object = inverse(trs_unit_scale(pos, euler(rot.x, rot.y, rot.z)));
  } else {
    cam_to_object = inverse(scene_cam);  // :266
  }
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 4; c++) f.cam[r * 4 + c] = cam_to_object.at(r, c);

  const float* bg = p.has_bg ? p.bg : s.bg;  // :322
  if (!p.has_bg && !s.has_image) { f.bg[0] = f.bg[1] = f.bg[2] = 0.2f; }
  else { f.bg[0] = bg[0]; f.bg[1] = bg[1]; f.bg[2] = bg[2]; }

  f.light[0] = f.light[1] = f.light[2] = 0.0f;  // :325-336: Lights[0] position = column 3 of its composite matrix
  if (s.n_lights > 0) {
    const int li = s.light_xforms[0];
    if (li >= 0 && li < s.n_xforms) {
      const Mat4 L = composite_matrix(s, li);
      f.light[0] = L.at(0, 3); f.light[1] = L.at(1, 3); f.light[2] = L.at(2, 3);
    }
  }
  const float fov = p.has_fov ? p.fov_deg : (s.has_camera ? s.cam_vfov_deg : 50.0f);  // :339
  f.cam_dist = s.has_camera ? s.cam_distance : 30.0f;                                   // :340
  // BVHRayTracing.compute:292: tan(radians(_CameraFOV) * 0.5); radians(x) = x * (pi/180) in FP32, tan through double.
  f.tan_half = (float)std::tan((double)((fov * 0.017453292f) * 0.5f));
  f.ortho_size = f.cam_dist * (float)std::tan((double)(kDeg2Rad * fov * 0.5f));  // RayTracer.cs:347
  f.spp = std::max(1, p.aa_samples);                                             // compute:283
  const float gs = std::sqrt((float)f.spp);
  f.grid_w = (int)std::ceil(gs);
  f.grid_h = (int)std::ceil((float)f.spp / (float)f.grid_w);
  f.max_depth = p.max_depth;
  f.en_ambient = p.enable_ambient; f.en_diffuse = p.enable_diffuse; f.en_specular = p.enable_specular;
  f.en_refraction = p.enable_refraction; f.ortho = p.is_orthographic;
  f.soft = p.soft_shadows; f.glossy = p.glossy; f.blur = p.motion_blur; f.debug = p.debug_mode;
  f.light_intensity = p.light_intensity; f.light_size = p.light_size; f.roughness = p.roughness; f.shutter = p.shutter_speed;
  f.srgb = p.srgb_encode;
  f.band_world = p.band_world > 1 ? p.band_world : 1;
  f.band_rank = f.band_world > 1 ? p.band_rank : 0;
  f.band_rows = p.band_rows > 0 ? p.band_rows : 32;
  if (f.band_rank < 0 || f.band_rank >= f.band_world) { err = "band_rank out of range"; return false; }
  if (f.band_rows % 4 != 0) { err = "band_rows must be a multiple of 4"; return false; }
  f.out_compact = p.out_layout == RTB_OUT_COMPACT ? 1 : 0;
  return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// Reference-shape BVH: spatial median split on the longest axis, leaves of <= 4 triangles or whatever a failed partition
// leaves, BFS numbering with sibling pairs adjacent (BVHBuilder.cs:76-238).
// ---------------------------------------------------------------------------------------------------------------------
namespace {
struct BuildNode {
  float mn[3], mx[3];
  int32_t start, count;   // count > 0 <=> leaf
  int32_t left, right;
  int32_t depth;
};
}  // namespace

void build_reference_bvh(const float* raw, int32_t n, RefBvh& out) {
  out.nodes.clear(); out.perm.clear(); out.max_leaf = 0; out.max_depth = 0;
  if (n <= 0) return;
  std::vector<int32_t> order((size_t)n);
  for (int32_t i = 0; i < n; i++) order[(size_t)i] = i;
  auto vertex = [&](int32_t tri, int k, int axis) { return raw[(size_t)tri * 12 + (size_t)k * 4 + (size_t)axis]; };
  auto centroid = [&](int32_t tri, int axis) { return raw[(size_t)tri * 12 + (size_t)axis * 4 + 3]; };

  std::vector<BuildNode> pool;
  pool.reserve((size_t)n);
  // Explicit work list instead of recursion; node ids in `pool` are arbitrary (the BFS below renumbers them), but the
  // in-place partition of a range must finish before its sub-ranges are partitioned, which a stack guarantees.
  struct Work { int32_t node; };
  std::vector<Work> todo;
  pool.push_back(BuildNode{{0, 0, 0}, {0, 0, 0}, 0, n, -1, -1, 0});
  todo.push_back(Work{0});
  while (!todo.empty()) {
    const int32_t me = todo.back().node;
    todo.pop_back();
    const int32_t start = pool[(size_t)me].start, count = pool[(size_t)me].count, depth = pool[(size_t)me].depth;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};  // AABB.Empty
    for (int32_t k = 0; k < count; k++) {
      const int32_t t = order[(size_t)(start + k)];
      for (int v = 0; v < 3; v++)
        for (int a = 0; a < 3; a++) {
          const float x = vertex(t, v, a);
          mn[a] = x < mn[a] ? x : mn[a];  // Mathf.Min / Mathf.Max
          mx[a] = x > mx[a] ? x : mx[a];
        }
    }
    std::memcpy(pool[(size_t)me].mn, mn, sizeof mn);
    std::memcpy(pool[(size_t)me].mx, mx, sizeof mx);
    out.max_depth = std::max(out.max_depth, depth);
    if (count <= 4) continue;  // MAX_TRIANGLES_PER_LEAF
    const float size[3] = {mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2]};
    int axis = 0;
    if (size[1] > size[0]) axis = 1;
    if (size[2] > size[axis]) axis = 2;
    const float pivot = (mn[axis] + mx[axis]) * 0.5f;  // bounds.Center[axis]
    int32_t i = start, j = start + count - 1;          // Partition, BVHBuilder.cs:160-183
    while (i <= j) {
      if (centroid(order[(size_t)i], axis) < pivot) i++;
      else { std::swap(order[(size_t)i], order[(size_t)j]); j--; }
    }
    if (i == start || i == start + count) continue;  // partition failed: stays a (possibly large) leaf
    const int32_t l = (int32_t)pool.size();
    pool.push_back(BuildNode{{0, 0, 0}, {0, 0, 0}, start, i - start, -1, -1, depth + 1});
    pool.push_back(BuildNode{{0, 0, 0}, {0, 0, 0}, i, start + count - i, -1, -1, depth + 1});
    pool[(size_t)me].left = l; pool[(size_t)me].right = l + 1; pool[(size_t)me].count = 0;
    todo.push_back(Work{l + 1});
    todo.push_back(Work{l});
  }

  // Flatten, BVHBuilder.cs:189-238
  out.nodes.assign(8, 0.0f);
  out.perm.reserve((size_t)n);
  std::deque<std::pair<int32_t, int32_t>> q;  // (pool id, output slot)
  q.emplace_back(0, 0);
  int32_t n_out = 1;
  while (!q.empty()) {
    const auto [id, slot] = q.front();
    q.pop_front();
    const BuildNode& b = pool[(size_t)id];
    int32_t left_or_first, cnt;
    if (b.count > 0) {
      cnt = b.count;
      left_or_first = (int32_t)out.perm.size();
      for (int32_t k = 0; k < b.count; k++) out.perm.push_back(order[(size_t)(b.start + k)]);
      out.max_leaf = std::max(out.max_leaf, b.count);
    } else {
      cnt = 0;
      left_or_first = n_out;
      n_out += 2;
      out.nodes.resize((size_t)n_out * 8, 0.0f);
      q.emplace_back(b.left, left_or_first);
      q.emplace_back(b.right, left_or_first + 1);
    }
    float* o = &out.nodes[(size_t)slot * 8];
    o[0] = b.mn[0]; o[1] = b.mn[1]; o[2] = b.mn[2]; std::memcpy(&o[3], &left_or_first, 4);
    o[4] = b.mx[0]; o[5] = b.mx[1]; o[6] = b.mx[2]; std::memcpy(&o[7], &cnt, 4);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Scene text format
// ---------------------------------------------------------------------------------------------------------------------
namespace {

struct LineReader {
  std::vector<std::string> lines;
  size_t i = 0;

  explicit LineReader(const char* text, size_t len) {  // File.ReadAllLines: \n, \r\n and \r all end a line
    size_t p = 0;
    while (p < len) {
      size_t q = p;
      while (q < len && text[q] != '\n' && text[q] != '\r') q++;
      lines.emplace_back(text + p, q - p);
      if (q >= len) break;
      p = (text[q] == '\r' && q + 1 < len && text[q + 1] == '\n') ? q + 2 : q + 1;
    }
  }
  static std::string clean(const std::string& raw) {  // Clean(), SceneService.cs:258-267
    std::string s = raw;
    const size_t c = s.find("//");
    if (c != std::string::npos) s.resize(c);
    size_t b = 0, e = s.size();
    while (b < e && std::isspace((unsigned char)s[b])) b++;
    while (e > b && std::isspace((unsigned char)s[e - 1])) e--;
    return s.substr(b, e - b);
  }
  bool done() const { return i >= lines.size(); }
  std::string take() {
    if (i >= lines.size()) throw std::runtime_error("scene file ends inside a segment");
    return clean(lines[i++]);
  }
  void skip_brace() {  // ExpectOpeningBrace / ExpectClosingBrace :280-301: skip blanks, consume one line (logged, not fatal)
    while (i < lines.size() && clean(lines[i]).empty()) i++;
    i++;
  }
};

bool same_word(const std::string& a, const char* b) {  // IsSegment :272-275
  const size_t n = std::strlen(b);
  if (a.size() != n) return false;
  for (size_t k = 0; k < n; k++)
    if (std::tolower((unsigned char)a[k]) != std::tolower((unsigned char)b[k])) return false;
  return true;
}

double to_double(const std::string& tok) {  // double.Parse(NumberStyles.Float, InvariantCulture)
  if (tok.empty()) throw std::runtime_error("expected a number, found an empty line");
  char* end = nullptr;
  const double v = std::strtod(tok.c_str(), &end);
  if (end == tok.c_str() || *end != '\0') throw std::runtime_error("malformed number '" + tok + "'");
  return v;
}

std::vector<std::string> tokens(const std::string& s) {  // Split(' ', '\t', RemoveEmptyEntries)
  std::vector<std::string> out;
  size_t p = 0;
  while (p < s.size()) {
    while (p < s.size() && (s[p] == ' ' || s[p] == '\t')) p++;
    size_t q = p;
    while (q < s.size() && s[q] != ' ' && s[q] != '\t') q++;
    if (q > p) out.push_back(s.substr(p, q - p));
    p = q;
  }
  return out;
}

std::vector<double> numbers(const std::string& line, size_t need) {
  std::vector<double> v;
  for (const auto& t : tokens(line)) v.push_back(to_double(t));
  if (v.size() < need) throw std::runtime_error("too few numbers on line '" + line + "'");
  return v;
}

}  // namespace

void parse_scene_text(const char* text, size_t len, HostScene& sc) {
  LineReader in(text, len);
  sc = HostScene();
  sc.xform_offsets.assign(1, 0);
  while (!in.done()) {
    const std::string head = in.take();
    if (head.empty()) continue;
    if (same_word(head, "Image")) {  // :45-64
      in.skip_brace();
      const auto res = numbers(in.take(), 2);
      const auto bg = numbers(in.take(), 3);
      in.skip_brace();
      sc.d.has_image = 1;
      sc.d.image_w = (int32_t)res[0]; sc.d.image_h = (int32_t)res[1];
      for (int k = 0; k < 3; k++) sc.d.bg[k] = (float)bg[(size_t)k];
    } else if (same_word(head, "Transformation")) {  // :65-116
      in.skip_brace();
      while (!in.done()) {
        const std::string line = in.take();
        if (line == "}") break;
        if (line.empty()) continue;
        const auto tok = tokens(line);
        if (tok.empty()) continue;
        auto arg = [&](size_t k) -> float {
          if (k >= tok.size()) throw std::runtime_error("transformation element '" + line + "' lacks arguments");
          return (float)to_double(tok[k]);
        };
        rtb_xform_elem e{};
        if (tok[0] == "T") { e.type = RTB_XF_T; e.x = arg(1); e.y = arg(2); e.z = arg(3); }
        else if (tok[0] == "S") { e.type = RTB_XF_S; e.x = arg(1); e.y = arg(2); e.z = arg(3); }
        else if (tok[0] == "Rx") { e.type = RTB_XF_RX; e.angle_deg = arg(1); }
        else if (tok[0] == "Ry") { e.type = RTB_XF_RY; e.angle_deg = arg(1); }
        else if (tok[0] == "Rz") { e.type = RTB_XF_RZ; e.angle_deg = arg(1); }
        else continue;  // unknown element keyword: ignored
        sc.xform_elems.push_back(e);
      }
      sc.xform_offsets.push_back((int32_t)sc.xform_elems.size());
    } else if (same_word(head, "Camera")) {  // :117-138
      in.skip_brace();
      const int t = (int)to_double(in.take());
      const double dist = to_double(in.take());
      const double fov = to_double(in.take());
      in.skip_brace();
      sc.d.has_camera = 1; sc.d.cam_xform = t; sc.d.cam_distance = (float)dist; sc.d.cam_vfov_deg = (float)fov;
    } else if (same_word(head, "Light")) {  // :139-157
      in.skip_brace();
      const int t = (int)to_double(in.take());
      const auto rgb = numbers(in.take(), 3);
      in.skip_brace();
      sc.light_xforms.push_back(t);
      for (int k = 0; k < 3; k++) sc.light_rgb.push_back((float)rgb[(size_t)k]);
    } else if (same_word(head, "Material")) {  // :158-180
      in.skip_brace();
      const auto col = numbers(in.take(), 3);
      const auto k = numbers(in.take(), 5);
      in.skip_brace();
      sc.materials.push_back(rtb_material{(float)col[0], (float)col[1], (float)col[2], (float)k[0], (float)k[1], (float)k[2],
                                          (float)k[3], (float)k[4]});
    } else if (same_word(head, "Triangles")) {  // :181-210
      in.skip_brace();
      rtb_mesh mesh{};
      mesh.xform = (int)to_double(in.take());
      mesh.first_tri = (int64_t)sc.triangles.size();
      while (!in.done()) {
        const std::string line = in.take();
        if (line == "}") break;
        if (line.empty()) continue;
        rtb_triangle t{};
        t.material = (int)to_double(line);
        float* dst[3] = {t.v0, t.v1, t.v2};
        for (int v = 0; v < 3; v++) {
          const auto xyz = numbers(in.take(), 3);
          for (int c = 0; c < 3; c++) dst[v][c] = (float)xyz[(size_t)c];
        }
        sc.triangles.push_back(t);
      }
      mesh.n_tris = (int64_t)sc.triangles.size() - mesh.first_tri;
      sc.meshes.push_back(mesh);
    } else if (same_word(head, "Sphere") || same_word(head, "Box")) {  // :211-238
      const bool sphere = same_word(head, "Sphere");
      in.skip_brace();
      const int t = (int)to_double(in.take());
      const int m = (int)to_double(in.take());
      in.skip_brace();
      (sphere ? sc.spheres : sc.boxes).push_back(rtb_prim{t, m});
    }
    // anything else: skipped line by line, like the reference
  }
  sc.d.n_triangles = (int64_t)sc.triangles.size();
  sc.relink();
}

}  // namespace rtb
