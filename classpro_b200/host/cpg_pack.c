/* cpg_pack.c -- ASCII -> 2-bit packing of read sequences (host side of the staging path). */
#include <string.h>
#include "classpro_gpu.h"

/* ASCII -> code, -1 for anything but upper-case ACGT.  Filled when the library is loaded, not on
   first use: cpg_pack_seq is called from several packing threads at once. */
static signed char code[256];

__attribute__((constructor)) static void cpg_pack_init(void)
{ memset(code,-1,sizeof(code));
  code['A'] = 0; code['C'] = 1; code['G'] = 2; code['T'] = 3;
}

int cpg_pack_seq(const char *seq, int32_t rlen, uint8_t *out)
{
  int bad = 0;
  int32_t i = 0;
  for (; i+4 <= rlen; i += 4)
    { int a = code[(unsigned char)seq[i]], b = code[(unsigned char)seq[i+1]],
          c = code[(unsigned char)seq[i+2]], d = code[(unsigned char)seq[i+3]];
      bad |= (a|b|c|d);
      out[i >> 2] = (uint8_t)((a & 3) | ((b & 3) << 2) | ((c & 3) << 4) | ((d & 3) << 6));
    }
  if (i < rlen)
    { unsigned v = 0;
      for (int k = 0; i+k < rlen; k++)
        { int a = code[(unsigned char)seq[i+k]];
          bad |= a;
          v |= (unsigned)(a & 3) << (2*k);
        }
      out[i >> 2] = (uint8_t)v;
    }
  return bad < 0 ? 1 : 0;
}
