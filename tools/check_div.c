// exhaustive check: q' = fma(fma(-d, q, a), r, q) with q = a * r, r = RN(1/d) equals a / d for every finite float a (d = 9, 3)
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <omp.h>
int main() {
  const float ds[2] = {9.0f, 3.0f};
  for (int t = 0; t < 2; ++t) {
    const float d = ds[t], r = 1.0f / d;
    uint64_t bad = 0; uint32_t first = 0;
    #pragma omp parallel for reduction(+:bad)
    for (int64_t i = 0; i < (1LL << 32); ++i) {
      uint32_t u = (uint32_t)i; float a; memcpy(&a, &u, 4);
      if (!isfinite(a)) continue;
      float q = a * r;
      float e = fmaf(d, q, -a);
      float q2 = fmaf(-e, r, q);
      float ref = a / d;
      if (memcmp(&q2, &ref, 4) != 0) { bad++; if (!first) first = u; uint32_t x1,x2,x3,x4; memcpy(&x1,&q,4);memcpy(&x2,&e,4);memcpy(&x3,&q2,4);memcpy(&x4,&ref,4); printf("u=%08x q=%08x e=%08x q2=%08x ref=%08x\n",u,x1,x2,x3,x4); }
    }
    printf("d=%g mismatches=%llu first=0x%08x\n", d, (unsigned long long)bad, first);
  }
  return 0;
}
