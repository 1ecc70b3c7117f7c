// Lock-free union-find on an int parent array: every link points a root at a SMALLER index, so the root of a
// set is its smallest member.  parent[i] < 0 marks an element that takes no part.
#pragma once
#include "common.cuh"

namespace cb200 {

__device__ __forceinline__ int uf_find(const int* parent, int i) {
  int p = parent[i];
  while (p != i) {
    i = p;
    p = parent[i];
  }
  return i;
}

__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }  // a > b: hang a under b
    const int old = atomicMin(parent + a, b);
    if (old == a) return;  // a was still a root: linked
    a = old;               // somebody re-rooted a meanwhile; retry from there
  }
}

}  // namespace cb200
