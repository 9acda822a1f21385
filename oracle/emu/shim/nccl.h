// nccl.h - the few NCCL types jmb200.cu names (the library itself is resolved with dlopen at run time); oracle/emu only
#pragma once
#include <cstddef>
#include "cuda_runtime.h"
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclDouble = 8 } ncclDataType_t;
typedef enum { ncclSum = 0 } ncclRedOp_t;
