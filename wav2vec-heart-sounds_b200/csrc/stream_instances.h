// Resampler instances (UP, DOWN, D, PS) of the row-streaming fused kernel: PS warps form a team that shares one staged
// block of frames, each warp computing 1/PS of the UP phases (integer up-sampling, DOWN == 1, takes the run form:
// one warp per block).
#pragma once
#define FZ_I8 8, 1, 15, 1
#define FZ_I8N 8, 1, 22, 1
#define FZ_I16 33, 16, 30, 4
#define FZ_I16N 33, 16, 36, 4
#define FZ_I32 33, 32, 46, 4
#define FZ_I32N 33, 32, 52, 4
