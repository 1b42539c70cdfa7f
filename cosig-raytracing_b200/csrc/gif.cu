// gif.cu — the batch caller behind the render path: the reference's GifGenerator (Assets/Services/GifGenerator.cs), SURVEY §8f-3.
//
//   k_palette            ConvertToIndexed, GifGenerator.cs:346-369, on the device: RGBA8 frame (row 0 = bottom, as the resolve
//                        kernel leaves it) -> one palette index per pixel, rows flipped to GIF order (top row first).  A
//                        streaming kernel (4 B in, 1 B out per pixel): a frame leaves the GPU as 1 byte per pixel instead of 4.
//   gif_lzw              LzwCompress, :411-501, on the host: same code stream as the reference's string-dictionary coder
//                        (9..12-bit codes, one leading clear code, dictionary frozen — never cleared — once 4096 codes exist),
//                        built on an open-addressing (prefix code, byte) table instead of Dictionary<string,int>.
//   GifFile              WriteGifHeader :190-198, GenerateColorTable :219-247, WriteLoopExtension :203-213,
//                        WriteFrameData :256-293, trailer :149.
//   rtb_gif_save_indexed SaveGifAsync :82-155 for frames that are already palette indices: frames are compressed on host
//                        threads (the reference's Parallel.For over frames, :123-130), written in order.
//
// The entry points that need a context (device quantisation, the fused render -> quantise -> readback -> LZW sweep) are in api.cu.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "gif.hpp"

namespace rtb {

namespace {

// (int)(channel * 5.99f) clamped to 0..5, channel = byte / 255f as Texture2D.GetPixels() hands it out (GifGenerator.cs:353-356);
// individually rounded FP32 division and product, like the C# expression
__device__ __forceinline__ int cube_level(unsigned b) {
  const int v = (int)__fmul_rn(__fdiv_rn((float)b, 255.0f), 5.99f);
  return v < 0 ? 0 : (v > 5 ? 5 : v);
}

constexpr int kBlock = 256;

__global__ void __launch_bounds__(kBlock) k_palette(const uchar4* __restrict__ src, int width, int height, uint8_t* __restrict__ dst) {
  __shared__ uint8_t level[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) level[i] = (uint8_t)cube_level((unsigned)i);
  __syncthreads();
  auto index_of = [&](uchar4 p) -> unsigned { return (unsigned)level[p.x] * 36u + (unsigned)level[p.y] * 6u + (unsigned)level[p.z]; };
  if ((width & 3) == 0) {  // four pixels per thread: one 16-byte load, one 4-byte store
    const int quads_per_row = width >> 2;
    const int64_t n = (int64_t)quads_per_row * height;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      const int y_out = (int)(i / quads_per_row), xq = (int)(i - (int64_t)y_out * quads_per_row);
      const int y_in = height - 1 - y_out;  // :360-366
      const uint4 raw = __ldcs(reinterpret_cast<const uint4*>(src + (size_t)y_in * width) + xq);
      const uchar4 p0 = *reinterpret_cast<const uchar4*>(&raw.x), p1 = *reinterpret_cast<const uchar4*>(&raw.y);
      const uchar4 p2 = *reinterpret_cast<const uchar4*>(&raw.z), p3 = *reinterpret_cast<const uchar4*>(&raw.w);
      const unsigned packed = index_of(p0) | (index_of(p1) << 8) | (index_of(p2) << 16) | (index_of(p3) << 24);
      reinterpret_cast<unsigned*>(dst + (size_t)y_out * width)[xq] = packed;
    }
  } else {
    const int64_t n = (int64_t)width * height;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      const int y_out = (int)(i / width), x = (int)(i - (int64_t)y_out * width);
      dst[i] = (uint8_t)index_of(src[(size_t)(height - 1 - y_out) * width + x]);
    }
  }
}

// ---- LZW ----------------------------------------------------------------------------------------------------------------
struct BitWriter {  // WriteCode, GifGenerator.cs:432-442: codes packed LSB first
  uint8_t* out;
  size_t n = 0;
  uint32_t buffer = 0;
  int count = 0;
  void put(int code, int size) {
    buffer |= (uint32_t)code << count;
    count += size;
    while (count >= 8) { out[n++] = (uint8_t)(buffer & 0xFF); buffer >>= 8; count -= 8; }
  }
  void flush() { if (count > 0) out[n++] = (uint8_t)buffer; }  // :495-498
};

// (prefix code, next byte) -> code.  The dictionary only ever grows to 4096 - 258 entries and is never reset, so a fixed
// open-addressing table of 8192 slots (load <= 0.47) does; key = prefix << 8 | byte, empty = 0xffffffff.
struct CodeTable {
  static constexpr int kSlots = 8192;
  uint32_t key[kSlots];
  uint16_t val[kSlots];
  CodeTable() { std::memset(key, 0xff, sizeof(key)); }
  static uint32_t hash(uint32_t k) { return (k * 2654435761u) >> 19; }  // top 13 bits
  int find(uint32_t k) const {
    for (uint32_t h = hash(k);; h = (h + 1) & (kSlots - 1)) {
      if (key[h] == k) return val[h];
      if (key[h] == 0xffffffffu) return -1;
    }
  }
  void insert(uint32_t k, int code) {
    uint32_t h = hash(k);
    while (key[h] != 0xffffffffu) h = (h + 1) & (kSlots - 1);
    key[h] = k; val[h] = (uint16_t)code;
  }
};

}  // namespace

void launch_palette(const void* rgba8, int width, int height, uint8_t* indexed, cudaStream_t st) {
  if (width <= 0 || height <= 0) return;
  const int64_t work = ((width & 3) == 0) ? (int64_t)(width >> 2) * height : (int64_t)width * height;
  int64_t grid = (work + kBlock - 1) / kBlock;
  if (grid > 148 * 16) grid = 148 * 16;
  k_palette<<<(int)grid, kBlock, 0, st>>>((const uchar4*)rgba8, width, height, indexed);
}

void gif_color_table(uint8_t* table) {  // :219-247
  int idx = 0;
  for (int r = 0; r < 6; r++)
    for (int g = 0; g < 6; g++)
      for (int b = 0; b < 6; b++) { table[idx++] = (uint8_t)(r * 51); table[idx++] = (uint8_t)(g * 51); table[idx++] = (uint8_t)(b * 51); }
  for (int i = 216; i < 256; i++) {
    const uint8_t gray = (uint8_t)(int)((float)(i - 216) * 6.5f);
    table[idx++] = gray; table[idx++] = gray; table[idx++] = gray;
  }
}

size_t gif_lzw_bound(size_t n) { return (n + 2) * 3 / 2 + 8; }  // <= 12 bits per input byte + clear + end codes

size_t gif_lzw(const uint8_t* data, size_t n, uint8_t* out) {
  const int clear_code = 256, end_code = 257;
  int next_code = end_code + 1, code_size = 9;
  BitWriter w{out};
  w.put(clear_code, code_size);  // :445
  if (n == 0) { w.put(end_code, code_size); w.flush(); return w.n; }
  CodeTable table;
  int current = data[0];  // codes 0..255 are the single bytes (:424-427)
  for (size_t i = 1; i < n; i++) {
    const uint32_t k = ((uint32_t)current << 8) | data[i];
    const int found = table.find(k);
    if (found >= 0) { current = found; continue; }  // :460-463
    w.put(current, code_size);                      // :468
    if (next_code < 4096) {                         // :471-481
      table.insert(k, next_code);
      if (next_code == (1 << code_size)) code_size++;
      next_code++;
    }
    current = data[i];
  }
  w.put(current, code_size);  // :489-490
  w.put(end_code, code_size);
  w.flush();
  return w.n;
}

namespace {
void put16(std::vector<uint8_t>& v, int x) { v.push_back((uint8_t)(x & 0xFF)); v.push_back((uint8_t)((x >> 8) & 0xFF)); }
}  // namespace

void gif_append_prologue(std::vector<uint8_t>& v, int width, int height) {
  const uint8_t sig[6] = {0x47, 0x49, 0x46, 0x38, 0x39, 0x61};  // :192
  v.insert(v.end(), sig, sig + 6);
  put16(v, width); put16(v, height);
  v.push_back(0xF7); v.push_back(0x00); v.push_back(0x00);
  uint8_t table[768];
  gif_color_table(table);
  v.insert(v.end(), table, table + 768);
  const uint8_t loop[19] = {0x21, 0xFF, 0x0B, 'N', 'E', 'T', 'S', 'C', 'A', 'P', 'E', '2', '.', '0', 0x03, 0x01, 0x00, 0x00, 0x00};  // :205-212
  v.insert(v.end(), loop, loop + 19);
}

void gif_append_frame(std::vector<uint8_t>& v, int width, int height, const uint8_t* compressed, size_t len, int delay_cs) {  // :256-293
  const uint8_t gce[4] = {0x21, 0xF9, 0x04, 0x00};
  v.insert(v.end(), gce, gce + 4);
  put16(v, delay_cs);
  v.push_back(0x00); v.push_back(0x00);
  v.push_back(0x2C);
  put16(v, 0); put16(v, 0); put16(v, width); put16(v, height);
  v.push_back(0x00);
  v.push_back(0x08);
  for (size_t off = 0; off < len;) {
    const size_t block = len - off < 255 ? len - off : 255;
    v.push_back((uint8_t)block);
    v.insert(v.end(), compressed + off, compressed + off + block);
    off += block;
  }
  v.push_back(0x00);
}

int gif_threads(int requested, int n_frames) {
  int t = requested > 0 ? requested : (int)std::thread::hardware_concurrency();
  if (t < 1) t = 1;
  if (t > n_frames) t = n_frames;
  return t < 1 ? 1 : t;
}

}  // namespace rtb

using namespace rtb;

extern "C" {

void rtb_gif_color_table(uint8_t* rgb768) { if (rgb768) gif_color_table(rgb768); }

int64_t rtb_gif_lzw_bound(int64_t n) { return n < 0 ? -1 : (int64_t)gif_lzw_bound((size_t)n); }

int64_t rtb_gif_lzw(const uint8_t* indexed, int64_t n, uint8_t* out, int64_t capacity) {
  if (n < 0 || !out || (n > 0 && !indexed)) return RTB_E_ARG;
  if (capacity < (int64_t)gif_lzw_bound((size_t)n)) return RTB_E_SIZE;
  return (int64_t)gif_lzw(indexed, (size_t)n, out);
}

int rtb_gif_save_indexed(const char* path, int32_t width, int32_t height, const uint8_t* const* frames, int32_t n_frames, int32_t frame_delay_cs,
                         int32_t threads) {
  if (!path || !frames || n_frames <= 0 || width <= 0 || height <= 0) return RTB_E_ARG;  // "frames == null || Count == 0 -> return", :84
  const size_t n_px = (size_t)width * height;
  std::vector<std::vector<uint8_t>> compressed((size_t)n_frames);
  std::atomic<int> next{0};
  auto work = [&] {
    for (int k = next.fetch_add(1); k < n_frames; k = next.fetch_add(1)) {
      compressed[(size_t)k].resize(gif_lzw_bound(n_px));
      compressed[(size_t)k].resize(gif_lzw(frames[k], n_px, compressed[(size_t)k].data()));
    }
  };
  const int t = gif_threads(threads, n_frames);
  std::vector<std::thread> pool;
  for (int i = 1; i < t; i++) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
  std::vector<uint8_t> file;
  gif_append_prologue(file, width, height);
  for (int k = 0; k < n_frames; k++) gif_append_frame(file, width, height, compressed[(size_t)k].data(), compressed[(size_t)k].size(), frame_delay_cs);
  file.push_back(0x3B);
  FILE* f = std::fopen(path, "wb");
  if (!f) return RTB_E_IO;
  const size_t wrote = std::fwrite(file.data(), 1, file.size(), f);
  const int rc = std::fclose(f);
  return (wrote == file.size() && rc == 0) ? RTB_OK : RTB_E_IO;
}

}  // extern "C"
