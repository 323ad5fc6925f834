/* tools/verify_constdiv.c — design-time proof (by exhaustion over all 2^23 significands) that the
 * 3-instruction division used on the device,
 *     q0 = x*rc;  r = fma(-q0, d, x);  q = fma(r, rc, q0)      with rc = RN(1/d),
 * returns exactly the IEEE-754 RN quotient x/d for every constant divisor of the hot path
 * (color_conversions.rs:158,168,177-179,184-187).  Scaling x by 2^k scales q0, r and q exactly, so one
 * binade of x covers all normal inputs whose quotient stays normal.
 * Build: gcc -O2 -ffp-contract=off -mfma tools/verify_constdiv.c -o /tmp/verify_constdiv -lm */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static long check(float d, const char *name) {
  float rc = 1.0f / d;
  long bad = 0;
  for (uint32_t m = 0; m < (1u << 23); m++) {
    uint32_t b = 0x3F800000u | m;
    float x; memcpy(&x, &b, 4);
    float q0 = x * rc;
    float r = fmaf(-q0, d, x);
    float q = fmaf(r, rc, q0);
    float ref = x / d;
    if (memcmp(&q, &ref, 4)) { if (bad < 3) printf("  %s: x=%a got %a want %a\n", name, x, q, ref); bad++; }
  }
  printf("%-12s d=%.9g rc=%a mismatches=%ld\n", name, d, rc, bad);
  return bad;
}
int main(void) {
  long bad = 0;
  float k = 24389.0f / 27.0f;
  bad += check(0.95047f, "white_x"); bad += check(1.08883f, "white_z"); bad += check(100.0f, "100");
  bad += check(255.0f, "255"); bad += check(116.0f, "116"); bad += check(500.0f, "500");
  bad += check(200.0f, "200"); bad += check(k, "k"); bad += check(65535.0f, "65535"); bad += check(12.92f, "12.92");
  bad += check(3.0f, "3"); bad += check(5.0f, "5"); bad += check(6.0f, "6"); bad += check(7.0f, "7"); bad += check(9.0f, "9");
  return bad != 0;
}
