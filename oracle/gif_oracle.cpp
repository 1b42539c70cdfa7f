// gif_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as oracle.cpp: only tests/, smoke() and bench legs may
// load it).
//
// CPU restatement of the reference's GIF writer, Assets/Services/GifGenerator.cs, statement by statement and with the
// reference's own data structures (a string-keyed dictionary for LZW, a growing byte list), so that it shares nothing with
// the product's coder (cosig-raytracing_b200/csrc/gif.cu: open-addressing prefix table, device palette kernel).
//
// PARITY UNPINNED as far as the reference goes (it ships no GIF fixtures and cannot run here), but this part is pinned by an
// INDEPENDENT decoder: tests/test_gif_cpu.py decodes the files written here with PIL's GIF reader and must get back exactly
// the palette indices and colour table — an LZW stream or container that deviated from GIF89a would not decode.
// Unity's Texture2D.GetPixels() is closed source: a channel is taken as byte / 255f in FP32 (the documented Color32 -> Color
// conversion), then `(int)(c * 5.99f)` as written at GifGenerator.cs:353-355.

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

typedef std::vector<uint8_t> Bytes;

void write_u16(Bytes& w, int v) { w.push_back((uint8_t)(v & 0xFF)); w.push_back((uint8_t)((v >> 8) & 0xFF)); }  // BinaryWriter.Write(ushort): little endian

// WriteGifHeader, GifGenerator.cs:190-198
void write_gif_header(Bytes& w, int width, int height) {
  const uint8_t sig[] = {0x47, 0x49, 0x46, 0x38, 0x39, 0x61};
  w.insert(w.end(), sig, sig + 6);
  write_u16(w, (uint16_t)width);
  write_u16(w, (uint16_t)height);
  w.push_back(0xF7);
  w.push_back(0x00);
  w.push_back(0x00);
}

// WriteLoopExtension, :203-213
void write_loop_extension(Bytes& w) {
  w.push_back(0x21);
  w.push_back(0xFF);
  w.push_back(0x0B);
  const char* app = "NETSCAPE2.0";
  for (int i = 0; i < 11; i++) w.push_back((uint8_t)app[i]);
  w.push_back(0x03);
  w.push_back(0x01);
  write_u16(w, 0x0000);
  w.push_back(0x00);
}

// GenerateColorTable, :219-247
Bytes generate_color_table() {
  Bytes table(256 * 3);
  int idx = 0;
  for (int r = 0; r < 6; r++) {
    for (int g = 0; g < 6; g++) {
      for (int b = 0; b < 6; b++) {
        table[idx++] = (uint8_t)(r * 51);
        table[idx++] = (uint8_t)(g * 51);
        table[idx++] = (uint8_t)(b * 51);
      }
    }
  }
  for (int i = 216; i < 256; i++) {
    const float scaled = (float)(i - 216) * 6.5f;
    const uint8_t gray = (uint8_t)(int)scaled;
    table[idx++] = gray;
    table[idx++] = gray;
    table[idx++] = gray;
  }
  return table;
}

// ConvertToIndexed, :346-369.  pixels: RGBA8, row 0 = bottom (Texture2D order).
Bytes convert_to_indexed(const uint8_t* rgba8, int width, int height) {
  const size_t n = (size_t)width * height;
  Bytes indexed(n);
  for (size_t i = 0; i < n; i++) {
    const float pr = (float)rgba8[4 * i] / 255.0f, pg = (float)rgba8[4 * i + 1] / 255.0f, pb = (float)rgba8[4 * i + 2] / 255.0f;
    const float fr = pr * 5.99f, fg = pg * 5.99f, fb = pb * 5.99f;
    const int r = std::max(0, std::min(5, (int)fr));
    const int g = std::max(0, std::min(5, (int)fg));
    const int b = std::max(0, std::min(5, (int)fb));
    indexed[i] = (uint8_t)(r * 36 + g * 6 + b);
  }
  Bytes flipped(n);
  for (int y = 0; y < height; y++)  // :360-366
    std::copy(indexed.begin() + (size_t)y * width, indexed.begin() + (size_t)(y + 1) * width, flipped.begin() + (size_t)(height - 1 - y) * width);
  return flipped;
}

// LzwCompress, :411-501
Bytes lzw_compress(const uint8_t* data, size_t length, int min_code_size) {
  Bytes output;
  const int clear_code = 1 << min_code_size;
  const int end_code = clear_code + 1;
  std::unordered_map<std::string, int> code_table;
  int next_code = end_code + 1;
  int code_size = min_code_size + 1;
  for (int i = 0; i < clear_code; i++) code_table[std::string(1, (char)i)] = i;
  int bit_buffer = 0;
  int bit_count = 0;
  auto write_code = [&](int code, int size) {
    bit_buffer |= code << bit_count;
    bit_count += size;
    while (bit_count >= 8) {
      output.push_back((uint8_t)(bit_buffer & 0xFF));
      bit_buffer >>= 8;
      bit_count -= 8;
    }
  };
  write_code(clear_code, code_size);
  if (length == 0) {
    write_code(end_code, code_size);
    if (bit_count > 0) output.push_back((uint8_t)bit_buffer);
    return output;
  }
  std::string current(1, (char)data[0]);
  for (size_t i = 1; i < length; i++) {
    const char c = (char)data[i];
    std::string next = current + c;
    if (code_table.count(next)) {
      current = next;
    } else {
      write_code(code_table[current], code_size);
      if (next_code < 4096) {
        code_table[next] = next_code;
        if (next_code == (1 << code_size)) code_size++;
        next_code++;
      }
      current = std::string(1, c);
    }
  }
  write_code(code_table[current], code_size);
  write_code(end_code, code_size);
  if (bit_count > 0) output.push_back((uint8_t)bit_buffer);
  return output;
}

// WriteFrameData, :256-293
void write_frame_data(Bytes& w, int width, int height, const Bytes& compressed, int delay) {
  w.push_back(0x21);
  w.push_back(0xF9);
  w.push_back(0x04);
  w.push_back(0x00);
  write_u16(w, (uint16_t)delay);
  w.push_back(0x00);
  w.push_back(0x00);
  w.push_back(0x2C);
  write_u16(w, 0);
  write_u16(w, 0);
  write_u16(w, (uint16_t)width);
  write_u16(w, (uint16_t)height);
  w.push_back(0x00);
  const uint8_t min_code_size = 8;
  w.push_back(min_code_size);
  size_t offset = 0;
  while (offset < compressed.size()) {
    const size_t block = std::min<size_t>(255, compressed.size() - offset);
    w.push_back((uint8_t)block);
    w.insert(w.end(), compressed.begin() + offset, compressed.begin() + offset + block);
    offset += block;
  }
  w.push_back(0x00);
}

}  // namespace

extern "C" {

void orc_gif_color_table(uint8_t* rgb768) {
  const Bytes t = generate_color_table();
  std::copy(t.begin(), t.end(), rgb768);
}

void orc_gif_convert_to_indexed(const uint8_t* rgba8, int32_t width, int32_t height, uint8_t* indexed) {
  const Bytes f = convert_to_indexed(rgba8, width, height);
  std::copy(f.begin(), f.end(), indexed);
}

// Returns the compressed length; writes at most `capacity` bytes.
int64_t orc_gif_lzw(const uint8_t* data, int64_t n, uint8_t* out, int64_t capacity) {
  const Bytes c = lzw_compress(data, (size_t)n, 8);
  if ((int64_t)c.size() <= capacity) std::copy(c.begin(), c.end(), out);
  return (int64_t)c.size();
}

// SaveGif, :160-184 (same bytes as SaveGifAsync :82-155): frames = n_frames RGBA8 images (row 0 = bottom), contiguous.
int orc_gif_save(const char* path, int32_t width, int32_t height, const uint8_t* rgba8_frames, int32_t n_frames, int32_t frame_delay) {
  if (n_frames <= 0) return 0;  // "frames == null || frames.Count == 0 -> return"
  Bytes w;
  write_gif_header(w, width, height);
  const Bytes table = generate_color_table();
  w.insert(w.end(), table.begin(), table.end());
  write_loop_extension(w);
  const size_t frame_bytes = (size_t)width * height * 4;
  for (int i = 0; i < n_frames; i++) {
    const Bytes indexed = convert_to_indexed(rgba8_frames + (size_t)i * frame_bytes, width, height);
    const Bytes compressed = lzw_compress(indexed.data(), indexed.size(), 8);
    write_frame_data(w, width, height, compressed, frame_delay);
  }
  w.push_back(0x3B);
  FILE* f = std::fopen(path, "wb");
  if (!f) return -1;
  const size_t wrote = std::fwrite(w.data(), 1, w.size(), f);
  std::fclose(f);
  return wrote == w.size() ? 0 : -1;
}

}  // extern "C"
