#pragma once
#include "common.cuh"
#include <vector>
namespace b200 {
// polyphase sinc table of torchaudio.functional.resample for orig:new (already divided by their gcd)
void resample_table(int orig, int nw, int* width_out, std::vector<float>* table);
// format 0 = int16 interleaved [frame][channel], 1 = float32 planar [channel][frame]; channel < 0 = downmix (mean)
int audio_ingest(const void* src, int format, int channels, long long frames_in, int channel, const float* table,
                 int orig, int nw, int width, float* out, long long frames_out, cudaStream_t stream);
}
