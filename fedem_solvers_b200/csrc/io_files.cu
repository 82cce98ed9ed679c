// io_files.cu -- host-side file formats and history assembly on the drop-in surface (no device code).
//
//  * tagged binary files of src/vpmUtilities/binaryDB.c (writeTagDB :643-671, readTagDB :676-733) with
//    the FFaTag header (fedem-foundation/src/FFaLib/FFaOS/FFaTag.C:192-297): 30-character tag,
//    16-bit endian mark 0x1234, 8-byte checksum field (4 zero bytes + the 32-bit checksum),
//    ";1.0;\n"  = 46 bytes, then raw arrays; files of the other endianness are byte-swapped on read;
//  * .fmx disk matrices (src/vpmUtilities/diskMatrixModule.f90:263-301): tag "#FEDEM disk matrix" or
//    "#FEDEM generalized modes", optional " SP" suffix = stored as float, column-major values; the
//    dimensions are NOT in the file (they come from the .fsm: ndof1 x ndof2 / ndof1 x ngen);
//  * .fsm SAM files (saveSAM, src/vpmReducer/samReducerModule.f90:586-672; read order of
//    readSAMarrays, src/vpmStress/samStressModule.f90:273-316): tag "#SAM data", npar, mpar(npar),
//    madof, minex, mnnn, msc, mpmnpc, mmnpc, melcon, [mpmceq, mmceq, ttcc], meqn, meqn1, meqn2;
//  * BuildFinit (src/vpmCommon/supElTypeModule.f90:1067-1114): the reduced displacement vector of a
//    step from the superelement and triad position matrices, batched over steps -> the Q the
//    recovery kernels consume.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "io_tagged.cuh"



using namespace fsr;

extern "C" {

int fsr_fmx_write(const char* path, const char* tag, int checksum, const double* A, long long n, int single_precision)
{
  if (!path || !A || n < 0) { set_error("fsr_fmx_write: bad arguments"); return FSR_ERR_ARG; }
  std::string tg = tag ? tag : "#FEDEM disk matrix";
  if (single_precision) tg += " SP";
  TaggedFile tf;
  int rc = tf.open_write(path, tg.c_str(), (unsigned int)checksum);
  if (rc) return rc;
  if (!single_precision) return tf.write(A, (size_t)n);
  std::vector<float> buf(1 << 16);
  for (long long i = 0; i < n; i += (long long)buf.size()) {
    const size_t m = (size_t)std::min<long long>((long long)buf.size(), n - i);
    for (size_t k = 0; k < m; ++k) buf[k] = (float)A[i + k];
    if ((rc = tf.write(buf.data(), m))) return rc;
  }
  return FSR_OK;
}

int fsr_fmx_read(const char* path, char* tag_out, int tag_cap, int* checksum, int* is_single, double* A, long long n)
{
  if (!path || n < 0 || (n > 0 && !A)) { set_error("fsr_fmx_read: bad arguments"); return FSR_ERR_ARG; }
  TaggedFile tf;
  int rc = tf.open_read(path);
  if (rc) return rc;
  std::string tg = tf.tag;
  bool sp = false;
  if (tg.size() >= 3 && tg.compare(tg.size() - 3, 3, " SP") == 0) { sp = true; tg.resize(tg.size() - 3); }
  if (tag_out && tag_cap > 0) { strncpy(tag_out, tg.c_str(), (size_t)tag_cap - 1); tag_out[tag_cap - 1] = 0; }
  if (checksum) *checksum = (int)tf.checksum;
  if (is_single) *is_single = sp ? 1 : 0;
  if (n == 0) return FSR_OK;
  if (!sp) return tf.read(A, (size_t)n, path);
  std::vector<float> buf(1 << 16);
  for (long long i = 0; i < n; i += (long long)buf.size()) {
    const size_t m = (size_t)std::min<long long>((long long)buf.size(), n - i);
    if ((rc = tf.read(buf.data(), m, path))) return rc;
    for (size_t k = 0; k < m; ++k) A[i + k] = (double)buf[k];
  }
  return FSR_OK;
}

int fsr_fsm_read_mpar(const char* path, int* checksum, int* mpar, int cap)
{
  if (!path || !mpar || cap < 1) { set_error("fsr_fsm_read_mpar: bad arguments"); return FSR_ERR_ARG; }
  TaggedFile tf;
  int rc = tf.open_read(path);
  if (rc) return rc;
  if (tf.tag != "#SAM data") { set_error("%s is not a SAM data file, tag=%s", path, tf.tag.c_str()); return FSR_ERR_ARG; }
  if (checksum) *checksum = (int)tf.checksum;
  int npar = 0;
  if ((rc = tf.read(&npar, 1, "npar"))) return rc;
  if (npar < 24 || npar > 10000) { set_error("%s: implausible MPAR size %d", path, npar); return FSR_ERR_ARG; }
  std::vector<int> mp((size_t)npar);
  if ((rc = tf.read(mp.data(), (size_t)npar, "mpar"))) return rc;
  for (int i = 0; i < cap; ++i) mpar[i] = i < npar ? mp[i] : 0;
  return npar;
}

int fsr_fsm_read(const char* path, int* madof, int* minex, int* mnnn, int* msc, int* mpmnpc, int* mmnpc, int* melcon,
                 int* mpmceq, int* mmceq, double* ttcc, int* meqn, int* meqn1, int* meqn2)
{
  if (!path || !madof || !msc || !mpmnpc || !mmnpc || !melcon || !meqn) { set_error("fsr_fsm_read: bad arguments"); return FSR_ERR_ARG; }
  TaggedFile tf;
  int rc = tf.open_read(path);
  if (rc) return rc;
  if (tf.tag != "#SAM data") { set_error("%s is not a SAM data file, tag=%s", path, tf.tag.c_str()); return FSR_ERR_ARG; }
  int npar = 0;
  if ((rc = tf.read(&npar, 1, "npar"))) return rc;
  if (npar < 24 || npar > 10000) { set_error("%s: implausible MPAR size %d", path, npar); return FSR_ERR_ARG; }
  std::vector<int> mp((size_t)npar);
  if ((rc = tf.read(mp.data(), (size_t)npar, "mpar"))) return rc;
  const int nnod = mp[0], nel = mp[1], ndof = mp[2], ndof1 = mp[3], ndof2 = mp[4], nceq = mp[6], nmmnpc = mp[14], nmmceq = mp[15];
  std::vector<int> skip;
  auto rd = [&](int* dst, size_t n, const char* what) -> int {
    if (dst) return tf.read(dst, n, what);
    skip.resize(n);
    return tf.read(skip.data(), n, what);
  };
  if ((rc = rd(madof, (size_t)nnod + 1, "madof"))) return rc;
  if ((rc = rd(minex, (size_t)nnod, "minex"))) return rc;
  if ((rc = rd(mnnn, (size_t)nnod, "mnnn"))) return rc;
  if ((rc = rd(msc, (size_t)ndof, "msc"))) return rc;
  if ((rc = rd(mpmnpc, (size_t)nel + 1, "mpmnpc"))) return rc;
  if ((rc = rd(mmnpc, (size_t)nmmnpc, "mmnpc"))) return rc;
  if ((rc = rd(melcon, (size_t)nel, "melcon"))) return rc;
  if (nceq > 0) {
    if ((rc = rd(mpmceq, (size_t)nceq + 1, "mpmceq"))) return rc;
    if ((rc = rd(mmceq, (size_t)nmmceq, "mmceq"))) return rc;
    if (ttcc) { if ((rc = tf.read(ttcc, (size_t)nmmceq, "ttcc"))) return rc; }
    else { std::vector<double> t((size_t)nmmceq); if ((rc = tf.read(t.data(), (size_t)nmmceq, "ttcc"))) return rc; }
  } else if (mpmceq)
    mpmceq[0] = 1;
  if ((rc = rd(meqn, (size_t)ndof, "meqn"))) return rc;
  if (ndof1 > 0 && (rc = rd(meqn1, (size_t)ndof1, "meqn1"))) return rc;
  if (ndof2 > 0 && (rc = rd(meqn2, (size_t)ndof2, "meqn2"))) return rc;
  return FSR_OK;
}

int fsr_fsm_write(const char* path, int checksum, int npar, const int* mpar, const int* madof, const int* minex,
                  const int* mnnn, const int* msc, const int* mpmnpc, const int* mmnpc, const int* melcon,
                  const int* mpmceq, const int* mmceq, const double* ttcc, const int* meqn, const int* meqn1,
                  const int* meqn2)
{
  if (!path || !mpar || npar < 24) { set_error("fsr_fsm_write: bad arguments"); return FSR_ERR_ARG; }
  const int nnod = mpar[0], nel = mpar[1], ndof = mpar[2], ndof1 = mpar[3], ndof2 = mpar[4], nceq = mpar[6], nmmnpc = mpar[14], nmmceq = mpar[15];
  TaggedFile tf;
  int rc = tf.open_write(path, "#SAM data", (unsigned int)checksum);
  if (rc) return rc;
  std::vector<int> zeros;
  auto wr = [&](const int* src, size_t n) -> int {
    if (src) return tf.write(src, n);
    zeros.assign(n, 0);
    return tf.write(zeros.data(), n);
  };
  if ((rc = tf.write(&npar, 1)) || (rc = tf.write(mpar, (size_t)npar))) return rc;
  if ((rc = wr(madof, (size_t)nnod + 1)) || (rc = wr(minex, (size_t)nnod)) || (rc = wr(mnnn, (size_t)nnod)) ||
      (rc = wr(msc, (size_t)ndof)) || (rc = wr(mpmnpc, (size_t)nel + 1)) || (rc = wr(mmnpc, (size_t)nmmnpc)) ||
      (rc = wr(melcon, (size_t)nel)))
    return rc;
  if (nceq > 0) {
    if ((rc = wr(mpmceq, (size_t)nceq + 1)) || (rc = wr(mmceq, (size_t)nmmceq))) return rc;
    if (!ttcc) { set_error("fsr_fsm_write: nceq > 0 needs ttcc"); return FSR_ERR_ARG; }
    if ((rc = tf.write(ttcc, (size_t)nmmceq))) return rc;
  }
  if ((rc = wr(meqn, (size_t)ndof))) return rc;
  if (ndof1 > 0 && (rc = wr(meqn1, (size_t)ndof1))) return rc;
  if (ndof2 > 0 && (rc = wr(meqn2, (size_t)ndof2))) return rc;
  return FSR_OK;
}

// BuildFinit for nsteps steps.  3x4 position matrices are column-major (12 doubles).
//  sup_tr   [nsteps][12]           superelement position  (sup%supTr)
//  triad_ur [nsteps][ntriads][12]  triad positions        (triads(i)%p%ur)
//  tr_undef [ntriads][12]          undeformed triad positions in the superelement system (sup%TrUndeformed)
//  ndofs / first_dof [ntriads]     triad DOF count and 1-based first superelement DOF
//  gen_ur   [nsteps][ngen], gen_first_dof 1-based;   Q [ldq x nsteps] column-major
int fsr_build_finit(int nsteps, int ntriads, const double* sup_tr, const double* triad_ur, const double* tr_undef,
                    const int* ndofs, const int* first_dof, int ngen, const double* gen_ur, int gen_first_dof,
                    double* Q, int ldq)
{
  if (nsteps < 0 || ntriads < 0 || !sup_tr || (ntriads > 0 && (!triad_ur || !tr_undef || !ndofs || !first_dof)) || !Q ||
      (ngen > 0 && !gen_ur)) { set_error("fsr_build_finit: bad arguments"); return FSR_ERR_ARG; }
  for (int s = 0; s < nsteps; ++s) {
    const double* a = sup_tr + 12 * (size_t)s;
    double* q = Q + (size_t)ldq * s;
    // invert34 (manipMatrixModule.f90:426-438): b(:,1:3) = a(:,1:3)^T, b(:,4) = -b(:,1:3) a(:,4)
    double b[12];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) b[i + 3 * j] = a[j + 3 * i];
    for (int i = 0; i < 3; ++i) b[i + 9] = -(b[i] * a[9] + b[i + 3] * a[10] + b[i + 6] * a[11]);
    for (int t = 0; t < ntriads; ++t) {
      const int n = ndofs[t];
      if (n < 3) continue;
      const double* u = triad_ur + 12 * ((size_t)s * ntriads + t);
      const double* T0 = tr_undef + 12 * (size_t)t;
      // urLocal = matmul34(invSupTr, ur) (manipMatrixModule.f90:287-299)
      double ul[12];
      for (int c = 0; c < 4; ++c)
        for (int i = 0; i < 3; ++i) {
          double v = b[i] * u[3 * c] + b[i + 3] * u[3 * c + 1] + b[i + 6] * u[3 * c + 2];
          ul[i + 3 * c] = c == 3 ? v + b[i + 9] : v;
        }
      const int j = first_dof[t] - 1;
      if (j < 0 || j + (n >= 6 ? 6 : 3) > ldq) { set_error("fsr_build_finit: triad %d DOF range outside Q", t + 1); return FSR_ERR_ARG; }
      for (int i = 0; i < 3; ++i) q[j + i] = ul[i + 9] - T0[i + 9];
      if (n >= 6) {
        // dR = urLocal(:,1:3) . TrUndeformed(:,1:3); rotations = (dR(3,2), dR(1,3), dR(2,1))
        auto dR = [&](int r, int c) { return ul[r] * T0[3 * c] + ul[r + 3] * T0[3 * c + 1] + ul[r + 6] * T0[3 * c + 2]; };
        q[j + 3] = dR(2, 1);
        q[j + 4] = dR(0, 2);
        q[j + 5] = dR(1, 0);
      }
    }
    if (ngen > 0) {
      if (gen_first_dof < 1 || gen_first_dof - 1 + ngen > ldq) { set_error("fsr_build_finit: generalized DOF range outside Q"); return FSR_ERR_ARG; }
      for (int k = 0; k < ngen; ++k) q[gen_first_dof - 1 + k] = gen_ur[(size_t)s * ngen + k];
    }
  }
  return FSR_OK;
}

// readSupElModes (src/vpmStress/modesRoutines.f90:121-203): the columns of Q for a mode-shape expansion (fedem_modes = the K1
// expansion with eigenvectors instead of a time history).  The solver stores the eigenvector components of every triad in global
// directions ("Eigenvectors|Mode n" of the Triad, nDOFs x ncomp values; ncomp = 2 for damped modes: real and imaginary part) and
// those of the component modes under the Part; per triad the translational and the rotational triple are turned into the part's
// system with invert34(supTr)(:,1:3) = supTr(:,1:3)^T, the generalized DOFs are copied.
//  sup_tr [12] column-major 3x4; triad_eig: concatenation over the triads of eigVec(nDOFs*ncomp) exactly as ffr_getData returns it
//  (component l of triad i starts at offset n*(l-1)); gen_eig [ngen*ncomp]; Q [ldq x ncomp] column-major: column l = eigFinit(:,l).
int fsr_build_mode_finit(int ntriads, const double* sup_tr, const int* ndofs, const int* first_dof, const double* triad_eig, int ngen,
                         int gen_first_dof, const double* gen_eig, int ncomp, double* Q, int ldq)
{
  if (ntriads < 0 || !sup_tr || (ntriads > 0 && (!ndofs || !first_dof || !triad_eig)) || !Q || ncomp < 1 || (ngen > 0 && !gen_eig)) {
    set_error("fsr_build_mode_finit: bad arguments");
    return FSR_ERR_ARG;
  }
  size_t off = 0;
  for (int t = 0; t < ntriads; ++t) {
    const int n = ndofs[t], k = first_dof[t] - 1;
    if (k < 0 || k + (n >= 6 ? 6 : n >= 3 ? 3 : 0) > ldq) { set_error("fsr_build_mode_finit: triad %d DOF range outside Q", t + 1); return FSR_ERR_ARG; }
    for (int l = 0; l < ncomp; ++l) {
      const double* e = triad_eig + off + (size_t)n * l;
      double* q = Q + (size_t)ldq * l + k;
      for (int h = 0; h < 2; ++h) {         // translations, then rotations
        if (n < 3 * (h + 1)) break;
        for (int i = 0; i < 3; ++i)         // tInv(i,:) = supTr(:,i)
          q[3 * h + i] = sup_tr[3 * i] * e[3 * h] + sup_tr[3 * i + 1] * e[3 * h + 1] + sup_tr[3 * i + 2] * e[3 * h + 2];
      }
    }
    off += (size_t)n * ncomp;
  }
  if (ngen > 0) {
    if (gen_first_dof < 1 || gen_first_dof - 1 + ngen > ldq) { set_error("fsr_build_mode_finit: generalized DOF range outside Q"); return FSR_ERR_ARG; }
    for (int l = 0; l < ncomp; ++l)
      for (int k = 0; k < ngen; ++k) Q[(size_t)ldq * l + gen_first_dof - 1 + k] = gen_eig[(size_t)ngen * l + k];
  }
  return FSR_OK;
}

}  // extern "C"
