// Resampler instances (UP, DOWN, D, FR, PS) of the fused kernel for the configured CTA size.
#pragma once
#if MPCG_FZ_THREADS == 1024
#define FZ_I8 8, 1, 15, 2, 1
#define FZ_I8N 8, 1, 22, 2, 1
#define FZ_I16 33, 16, 30, 1, 4
#define FZ_I16N 33, 16, 36, 1, 4
#define FZ_I32 33, 32, 46, 1, 8
#define FZ_I32N 33, 32, 52, 1, 8
#elif MPCG_FZ_THREADS == 512
#define FZ_I8 8, 1, 15, 4, 1
#define FZ_I8N 8, 1, 22, 4, 1
#define FZ_I16 33, 16, 30, 1, 2
#define FZ_I16N 33, 16, 36, 1, 2
#define FZ_I32 33, 32, 46, 1, 4
#define FZ_I32N 33, 32, 52, 1, 4
#else
#error "MPCG_FZ_THREADS must be 512 or 1024"
#endif
