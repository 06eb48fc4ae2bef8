#include "drfe_internal.h"
#include <cstdio>
int main() {
  int start[drfe::ChunkPipe::kMaxChunks + 1];
  for (int wish : {0, -1, 1, 5, 32, 64, 1000})
    for (int n = 1; n <= 3000; n += (n < 300 ? 1 : 37)) {
      const int c = drfe::ChunkPipe::schedule(n, wish, start);
      if (c < 1 || c > drfe::ChunkPipe::kMaxChunks || start[0] != 0 || start[c] != n) { printf("bad %d %d\n", wish, n); return 1; }
      for (int k = 0; k < c; ++k) if (start[k + 1] <= start[k]) { printf("empty chunk %d %d %d\n", wish, n, k); return 1; }
    }
  const int c = drfe::ChunkPipe::schedule(256, 0, start);
  printf("%d:", c);
  for (int k = 0; k < c; ++k) printf(" %d", start[k + 1] - start[k]);
  printf("\n");
  return 0;
}
