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

/* Class string of a read from its packed interval table (include/classpro_gpu.h, "compact results"):
   what the reference writes into rasgn, src/ClassPro.c:114-117 ('N' x (K-1)) and :265-271. */
void cpg_expand_intervals(int32_t K, int32_t rlen, const uint32_t *ivl, int32_t n, char *out)
{ static const char cls_chr[8] = { 'E','R','H','D','?','?','?','?' };
  int32_t b = 0;
  const int32_t plen = rlen-K+1;
  memset(out,'N',(size_t)(K-1 < rlen ? K-1 : rlen));
  if (plen <= 0) return;
  out += K-1;
  for (int32_t i = 0; i < n; i++)
    { int32_t e = (int32_t)(ivl[i] >> 3);
      if (e > plen) e = plen;
      if (e > b) { memset(out+b,cls_chr[ivl[i] & 7u],(size_t)(e-b)); b = e; }
    }
}
