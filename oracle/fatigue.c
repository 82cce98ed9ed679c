/* fatigue.c -- oracle (TEST INFRASTRUCTURE, see oracle.h): peak-valley extraction, rainflow
 * counting, S-N damage.  Follows fedem-foundation/src/FFpLib/FFpFatigue/FFpFatigue.C:77-163
 * (FFpPVXprocessor::process / locateFirstTP), :185-320 (FFpRainFlowCycleCounter),
 * :381-396 (getDamage), FFpSNCurve.C:10-33 (NorSok curve), FFpCycle.C (range),
 * FFpFatigue_F.C:81-141 (ffp_getdamage / ffp_getnumcycles call sequence).
 * Checked cycle-for-cycle against the reference's own compiled C++ (oracle/_ref) in tests/. */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>

/* FFpFatigue.C:129-163.  Returns index of first turning point, or -1-iTP if none. */
static int locate_first_tp(const double *data, int nData, double gate, double *possibleTP,
                           double *deltaOut)
{
  double deltaTP = data[0];
  int iMin = 0, iMax = 0, iTP = 0;
  for (int i = 1; i < nData; i++)
    if ((data[i] - data[iTP]) * deltaTP > 0.0)
      iTP = i;
    else if (data[i - 1] - data[iMin] > gate)
      return iMin;
    else if (data[iMax] - data[i - 1] > gate)
      return iMax;
    else if (data[iTP] > data[iMax]) {
      iMax = iTP;
      deltaTP = data[i] - data[iTP];
    } else if (data[iTP] < data[iMin]) {
      iMin = iTP;
      deltaTP = data[i] - data[iTP];
    } else {
      iTP = i;
      deltaTP = data[i] - data[iTP];
    }
  iTP = iMax > iMin ? iMax : iMin;
  *possibleTP = data[iTP];
  *deltaOut = deltaTP;
  return -1 - iTP;
}

/* FFpFatigue.C:77-126 with isFirstData = isLastData = true (the ffp_getdamage call).
 * turns must hold n doubles.  Returns the number of turning points. */
int orc_pvx(const double *data, int n, double gate, double *turns)
{
  int nt = 0, iFirst = 0;
  double possibleTP = 0.0, deltaTP = 0.0;
  if (n > 0) {
    iFirst = locate_first_tp(data, n, gate, &possibleTP, &deltaTP);
    if (iFirst < 0) return 0;
    possibleTP = data[iFirst];
    deltaTP = data[iFirst + 1] - data[iFirst];
    turns[nt++] = possibleTP;
    ++iFirst;
  }
  for (int i = iFirst; i < n; i++) {
    double delta = data[i] - possibleTP;
    if (delta * deltaTP <= 0.0) {
      if (fabs(delta) > gate)
        turns[nt++] = possibleTP;
      else
        continue;
    }
    possibleTP = data[i];
    deltaTP = data[i] - (nt == 0 ? 0.0 : turns[nt - 1]);
  }
  if (fabs(deltaTP) > gate) turns[nt++] = possibleTP;
  return nt;
}

/* std::list<double> emulated with index links; node 0 is the end() sentinel. */
typedef struct {
  double *val;
  int *next, *prev;
  int size, nalloc;
} tplist;

static int l_begin(const tplist *l) { return l->next[0]; }
static int l_erase(tplist *l, int it) /* returns iterator following the erased element */
{
  int n = l->next[it], p = l->prev[it];
  l->next[p] = n;
  l->prev[n] = p;
  l->size--;
  return n;
}
static int l_insert_before(tplist *l, int pos, double v)
{
  int id = l->nalloc++;
  int p = l->prev[pos];
  l->val[id] = v;
  l->next[id] = pos;
  l->prev[id] = p;
  l->next[p] = id;
  l->prev[pos] = id;
  l->size++;
  return id;
}

/* FFpFatigue.C:201-271 */
static int process_tp_list(tplist *l, double gate, double *cf, double *cs, int *ncyc)
{
  int tp[4], nRemoved = 0;
  double range[3];
  if (l->size < 4) return 0;
  tp[0] = l_begin(l);
  tp[1] = l->next[tp[0]];
  tp[2] = l->next[tp[1]];
  tp[3] = l->next[tp[2]];
  while (tp[3] != 0) {
    for (int i = 0; i < 3; i++) range[i] = l->val[tp[i + 1]] - l->val[tp[i]];
    if (range[0] * range[1] > 0.0) {
      nRemoved++;
      tp[1] = l_erase(l, tp[1]);
      tp[2] = l->next[tp[2]];
      tp[3] = l->next[tp[3]];
    } else if (range[1] * range[2] > 0.0) {
      nRemoved++;
      tp[2] = l_erase(l, tp[2]);
      tp[3] = l->next[tp[3]];
    } else if (fabs(range[0]) >= fabs(range[1]) && fabs(range[2]) >= fabs(range[1])) {
      nRemoved += 2;
      if (fabs(range[1]) > gate) {
        cf[*ncyc] = l->val[tp[1]];
        cs[*ncyc] = l->val[tp[2]];
        (*ncyc)++;
      }
      l_erase(l, tp[1]);
      tp[1] = l_erase(l, tp[2]);
      tp[3] = l->next[tp[3]];
      if (tp[3] != 0) {
        tp[2] = tp[3];
        tp[3] = l->next[tp[3]];
      }
    } else
      for (int i = 0; i < 4; i++) tp[i] = l->next[tp[i]];
  }
  return nRemoved > 0;
}

/* FFpFatigue.C:185-198 + :274-320 with isLastData = true.  cyc_first/cyc_second must hold
 * nturns/2+2 entries.  Returns the number of counted cycles, or -1 if the closing step does
 * not end with exactly three points (the reference's failure return). */
int orc_rainflow(const double *turns, int nturns, double gate, double *cyc_first,
                 double *cyc_second)
{
  tplist l;
  int ncyc = 0, cap = 2 * nturns + 8;
  l.val = (double *)malloc(sizeof(double) * cap);
  l.next = (int *)malloc(sizeof(int) * cap);
  l.prev = (int *)malloc(sizeof(int) * cap);
  l.size = 0;
  l.nalloc = 1;
  l.next[0] = l.prev[0] = 0;
  for (int i = 0; i < nturns; i++) l_insert_before(&l, 0, turns[i]);

  while (process_tp_list(&l, gate, cyc_first, cyc_second, &ncyc))
    ;

  /* processFinish */
  if (l.size > 1) {
    int pos = l_begin(&l), maxPos = pos;
    for (pos = l.next[pos]; pos != 0; pos = l.next[pos])
      if (fabs(l.val[pos]) > fabs(l.val[maxPos])) maxPos = pos;
    l_insert_before(&l, maxPos, l.val[maxPos]);
    /* insert copies of [maxPos,end) at the beginning, then erase [maxPos,end) */
    {
      int first = l_begin(&l);
      for (int it = maxPos; it != 0; it = l.next[it]) l_insert_before(&l, first, l.val[it]);
      for (int it = maxPos; it != 0;) it = l_erase(&l, it);
    }
    while (process_tp_list(&l, gate, cyc_first, cyc_second, &ncyc))
      ;
    if (l.size != 3) ncyc = -1;
    else {
      int a = l_begin(&l);
      cyc_first[ncyc] = l.val[a];
      cyc_second[ncyc] = l.val[l.next[a]];
      ncyc++;
    }
  }
  free(l.val); free(l.next); free(l.prev);
  return ncyc;
}

/* FFpSNCurve.C:10-33, two-slope NorSok curve */
double orc_sn_norsok(double s, double loga1, double loga2, double m1, double m2)
{
  double logN0 = (m2 * loga1 - m1 * loga2) / (m2 - m1);
  double logN = loga1 - m1 * log10(s);
  if (logN < logN0) return pow(10.0, logN);
  logN = loga2 - m2 * log10(s);
  return pow(10.0, logN);
}

static int cmp_double(const void *a, const void *b)
{
  double x = *(const double *)a, y = *(const double *)b;
  return (x > y) - (x < y);
}

/* FFpFatigue_F.C:99-123 + FFpFatigue.C:381-396: cycles sorted by range (toMPaScale = 1),
 * Miner sum in that order. */
double orc_damage(const double *cyc_first, const double *cyc_second, int ncyc, double loga1,
                  double loga2, double m1, double m2)
{
  double damage = 0.0;
  double *r;
  if (ncyc <= 0) return 0.0;
  r = (double *)malloc(sizeof(double) * ncyc);
  for (int i = 0; i < ncyc; i++) r[i] = fabs(cyc_first[i] - cyc_second[i]);
  qsort(r, ncyc, sizeof(double), cmp_double);
  for (int i = 0; i < ncyc; i++) damage += 1.0 / orc_sn_norsok(r[i], loga1, loga2, m1, m2);
  free(r);
  return damage;
}
