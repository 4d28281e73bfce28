// dist_bits.hpp — host only, no CUDA: index logic of a global<->local qubit exchange.
#pragma once

// pure index logic of a remap, shared by the GPU path and the host-only planner introspection:
// physical bit b in [nl-g, nl) <-> b + g
inline int dist_swap_bit(int b, int nl, int g) {
    if (b >= nl - g && b < nl) return b + g;
    if (b >= nl && b < nl + g) return b - g;
    return b;
}
