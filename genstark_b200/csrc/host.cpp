// Host-only translation unit (g++): AIR parsing and execution-trace generation.
#define GS_HOSTAIR_IMPL
#include "hostair.h"
