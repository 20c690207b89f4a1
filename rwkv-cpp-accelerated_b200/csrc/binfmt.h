// binfmt.h — forwarding header: the format table lives in include/rwkv/rwkv/format.h so that
// the public host API and the engine share one definition.
#pragma once
#include "../../include/rwkv/rwkv/format.h"
namespace binfmt = rwkv_format;
