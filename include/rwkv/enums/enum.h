// Tensor-table indices and forward modes of the RWKV-v4 uint8 ".bin" format.
//
// API parity: the enumerator names and their numeric order are the public contract
// of the reference (include/rwkv/enums/enum.h:2-54) because the order is also the
// on-disk order of the model file (converter/cpp_save_tensor.cpp:79-93) and the
// index into RWKV::tensors. Shapes below use L = n_layers, E = n_embed, V = 50277.
#pragma once

enum MODE
{
    PARRALEL, // T tokens = T independent streams, state slot t per token (sic: reference spelling)
    GPT       // T tokens = one stream processed in order, state slot 0
};

enum
{
    X = 0,           // f64 [E]        residual-stream scratch
    EMBED = 1,       // f32 [V][E]     embedding table
    LAYERNORMS = 2,  // f64 [4(L+1)][E] ln0 w,b | per layer ln1 w,b, ln2 w,b | ln_out w,b
    STATEXY = 3,     // f64 [L][E]     token-shift state of the attention block
    STATEAA = 4,     // f64 [L][E]     WKV numerator state
    STATEBB = 5,     // f64 [L][E]     WKV denominator state
    STATEPP = 6,     // f64 [L][E]     carried but never changed by the CUDA forward
    STATEDD = 7,     // f64 [L][E]     token-shift state of the FFN block
    BUFFER1 = 8,     // f64 [E]        scratch
    BUFFER2 = 9,     // f32 [V]        scratch / logits
    BUFFER3 = 10,    // f32 [E]        scratch
    BUFFER4 = 11,    // f32 [E]        scratch
    MIXK = 12,       // f64 [L][E]     att time_mix_k
    MIXV = 13,       // f64 [L][E]     att time_mix_v
    MIXR = 14,       // f64 [L][E]     att time_mix_r
    KM = 15,         // u8  [L][E][E]  att key weights, [in][out]
    VM = 16,         // u8  [L][E][E]  att value weights
    RM = 17,         // u8  [L][E][E]  att receptance weights
    KR = 18,         // f32 [L][E]     per-input-row scale of KM
    VR = 19,         // f32 [L][E]
    RR = 20,         // f32 [L][E]
    O1 = 21,         // f32 [L][E]     per-input-row offset of KM
    O2 = 22,         // f32 [L][E]     ... of VM
    O3 = 23,         // f32 [L][E]     ... of RM
    ATTOUT = 24,     // u8  [L][E][E]  att output projection
    ATTOUTR = 25,    // f32 [L][E]
    ATTOUTO = 26,    // f32 [L][E]
    FFNMIXK = 27,    // f64 [L][E]     ffn time_mix_k
    FFNMIXV = 28,    // f64 [L][E]     ffn time_mix_r (named "v" by the reference)
    FFNK = 29,       // u8  [L][E][4E] ffn key
    FFNV = 30,       // u8  [L][4E][E] ffn value
    FFNR = 31,       // u8  [L][E][E]  ffn receptance
    FFNKR = 32,      // f32 [L][E]
    FFNVR = 33,      // f32 [L][4E]
    FFNRR = 34,      // f32 [L][E]
    FFNKO = 35,      // f32 [L][E]
    FFNVO = 36,      // f32 [L][4E]
    FFNRO = 37,      // f32 [L][E]
    FFNKBUFFER = 38, // f64 [E]        scratch
    FFNVBUFFER = 39, // f64 [E]        scratch
    FFNRBUFFER = 40, // f32 [4E]       scratch
    DECAY = 41,      // f64 [L][E]     -exp(time_decay)
    BONUS = 42,      // f64 [L][E]     time_first
    HEAD = 43,       // u8  [E][V]     output head
    HEADR = 44,      // f32 [E]
    HEADO = 45       // f32 [E]
};
