/* Writes NL lines '<0|1> R a<i> b<j> c<k>' of distinct cells of a 100000^3 cube (BASELINE config 4's shape) to argv[1].
   gcc -O2 -o gen_obs gen_obs.c ; ./gen_obs out.obs 100000000 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

static char *put_u(char *p, uint64_t v) {
    char t[24]; int n = 0;
    do { t[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = t[--n];
    return p;
}
int main(int argc, char **argv) {
    if (argc < 3) return 1;
    FILE *f = fopen(argv[1], "wb");
    const uint64_t nl = strtoull(argv[2], 0, 10), cube = 1000000000000000ull;
    static char buf[1 << 22];
    char *p = buf;
    uint64_t x = 88172645463325252ull;
    for (uint64_t l = 0; l < nl; l++) {
        const uint64_t idx = (l * 7777777777ull + 12345ull) % cube;   /* l -> idx is injective for l < cube: gcd(7777777777, 1e15) = 1 */
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        *p++ = (char)('0' + (x & 1)); *p++ = ' '; *p++ = 'R'; *p++ = ' ';
        *p++ = 'a'; p = put_u(p, idx / 10000000000ull); *p++ = ' ';
        *p++ = 'b'; p = put_u(p, (idx / 100000ull) % 100000ull); *p++ = ' ';
        *p++ = 'c'; p = put_u(p, idx % 100000ull); *p++ = '\n';
        if (p - buf > (1 << 22) - 64) { fwrite(buf, 1, (size_t)(p - buf), f); p = buf; }
    }
    fwrite(buf, 1, (size_t)(p - buf), f);
    fclose(f);
    return 0;
}
