// gif.hpp — internal interface of gif.cu (GIF sweep: palette quantisation kernel, LZW coder, container writer).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <vector>

#include "../../include/rtb.h"

namespace rtb {

// ConvertToIndexed (GifGenerator.cs:346-369) on the device: rgba8 = RGBA8 frame with row 0 = bottom; indexed = width*height
// palette indices, top row first.  Asynchronous on `st`.
void launch_palette(const void* rgba8, int width, int height, uint8_t* indexed, cudaStream_t st);
void gif_color_table(uint8_t* rgb768);
size_t gif_lzw_bound(size_t n);
size_t gif_lzw(const uint8_t* data, size_t n, uint8_t* out);  // LzwCompress, :411-501; returns bytes written
void gif_append_prologue(std::vector<uint8_t>& file, int width, int height);  // header + colour table + loop extension
void gif_append_frame(std::vector<uint8_t>& file, int width, int height, const uint8_t* compressed, size_t len, int delay_cs);
int gif_threads(int requested, int n_frames);

}  // namespace rtb
