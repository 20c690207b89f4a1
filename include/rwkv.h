// rwkv.h — the single include every consumer of the reference uses (include/rwkv.h:1-3 there):
// the RWKV / RWKVState host API, the typical() sampler and the GPT2Tokenizer.
#include "rwkv/rwkv/rwkv.h"
#include "rwkv/sampler/typical.h"
#include "rwkv/tokenizer/tokenizer.h"
