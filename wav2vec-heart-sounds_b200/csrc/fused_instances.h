// Resampler instances (UP, DOWN, D, PS) of the fused preprocess kernel: PS warps form a team that shares one staged
// block of 32 frames, each warp computing 1/PS of the UP phases.
#pragma once
#define FZ_I8 8, 1, 15, 1
#define FZ_I8N 8, 1, 22, 1
#define FZ_I16 33, 16, 30, 4
#define FZ_I16N 33, 16, 36, 4
#if defined(MPCG_FZ_THREADS) && MPCG_FZ_THREADS == 1024      /* 32 warps: eight-warp teams keep the staging inside its budget */
#define FZ_I32 33, 32, 46, 8
#define FZ_I32N 33, 32, 52, 8
#else
#define FZ_I32 33, 32, 46, 4
#define FZ_I32N 33, 32, 52, 4
#endif
