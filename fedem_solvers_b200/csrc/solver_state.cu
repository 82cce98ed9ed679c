// solver_state.cu -- the in-core recovery state fedempy polls during a dynamics run (host bookkeeping around
// the device calls).
//
// In the reference the dynamics solver keeps, per FE part with recovery switched on, the expanded displacements
// sv and the von Mises array vms of the current step in core (part(:) of src/vpmSolver/stressRecoveryModule.f90:
// 60-280, filled by the per-step stress recovery :517-768,991-1225), and exports
//   getPartDeformationStateSize / getPartStressStateSize / savePartDeformationState / savePartStressState
// (src/vpmSolver/solverInterface.C:940-1001 -> slv_partsize / slv_savepart, solverDriver.f90:435-485 ->
// partStateVectorSize / savePartState, solverModule.f90:2183-2263 -> getDeformation / getStress), which
// fedempy's FedemSolver.save_part_state() calls (PythonAPI/src/fedempy/solver.py:524-629).
// Here the same four entry points, under the same names, serve the parts registered with
// fsr_recovery_register; fsr_recovery_update plays the solver's per-step recovery (expansion + von Mises on
// the device for the step's reduced displacements).
#include <algorithm>
#include <map>
#include <vector>

#include "common.cuh"

using namespace fsr;

namespace {
struct Entry {
  fsr_part* part = nullptr;
  std::vector<int> minex;
  std::vector<double> sv, vms;
  double *sv_pin = nullptr, *vm_pin = nullptr;   // page-locked landing buffers of the step's device results
  double step = 0.0, time = 0.0, dt = 0.0;
  bool have_state = false;
};
std::map<int, Entry> g_parts;
}  // namespace

extern "C" {

int fsr_recovery_register(int base_id, fsr_part* part, const int* minex)
{
  if (!part || base_id < 1) { set_error("fsr_recovery_register: bad arguments"); return FSR_ERR_ARG; }
  Entry e;
  e.part = part;
  if (minex) e.minex.assign(minex, minex + part->nnod);
  e.sv.assign((size_t)part->ndof, 0.0);
  const int nvms = fsr_vms_size(part);
  e.vms.assign((size_t)(nvms > 0 ? nvms : 0), 0.0);
  fsr_recovery_unregister(base_id);
  FSR_CUDA(cudaSetDevice(part->device));
  FSR_CUDA(cudaMallocHost(&e.sv_pin, sizeof(double) * (size_t)std::max(part->ndof, 1)));
  FSR_CUDA(cudaMallocHost(&e.vm_pin, sizeof(double) * (size_t)std::max(part->npts, 1)));
  g_parts[base_id] = e;
  return FSR_OK;
}

int fsr_recovery_unregister(int base_id)
{
  auto it = g_parts.find(base_id);
  if (it == g_parts.end()) return FSR_ERR_ARG;
  if (it->second.sv_pin) cudaFreeHost(it->second.sv_pin);
  if (it->second.vm_pin) cudaFreeHost(it->second.vm_pin);
  g_parts.erase(it);
  return FSR_OK;
}

// The solver's recovery of one converged step for SEVERAL parts (the loop over the parts of
// stressRecoveryModule.f90:1021-1061): the expansion and the stress kernels of all parts are queued on their own streams /
// devices first and waited for afterwards, so the parts of a mechanism overlap instead of running one after the other.
// q[k] = [finit; vg] of part base_ids[k].  The parts' running von Mises envelopes take the step along.
int fsr_recovery_update_parts(int nparts, const int* base_ids, int step, double time, double time_step, const double* const* q)
{
  if (nparts < 0 || (nparts > 0 && (!base_ids || !q))) { set_error("fsr_recovery_update_parts: bad arguments"); return FSR_ERR_ARG; }
  std::vector<Entry*> es((size_t)nparts);
  for (int k = 0; k < nparts; ++k) {
    auto it = g_parts.find(base_ids[k]);
    if (it == g_parts.end() || !q[k]) { set_error("fsr_recovery_update: unknown part %d", base_ids[k]); return FSR_ERR_ARG; }
    es[(size_t)k] = &it->second;
  }
  for (int k = 0; k < nparts; ++k) {
    const int rc = step_enqueue(es[(size_t)k]->part, q[k], es[(size_t)k]->sv_pin, es[(size_t)k]->vm_pin);
    if (rc < 0) return rc;
  }
  for (int k = 0; k < nparts; ++k) {
    Entry& e = *es[(size_t)k];
    const fsr_part* p = e.part;
    const int rc = fsr_synchronize(e.part);
    if (rc < 0) return rc;
    std::copy(e.sv_pin, e.sv_pin + p->ndof, e.sv.begin());
    // the in-core vms layout (stressRoutines.f90:324-331): [iel, nenod, nstrp, vm(1..nstrp)] per element with stress points
    size_t m = 0;
    for (int iel = 0; iel < p->nel; ++iel) {
      const int nstrp = p->ptoff_host[(size_t)iel + 1] - p->ptoff_host[(size_t)iel];
      if (nstrp <= 0) continue;
      e.vms[m++] = (double)(iel + 1);
      e.vms[m++] = (double)p->nenod_host[(size_t)iel];
      e.vms[m++] = (double)nstrp;
      for (int i = 0; i < nstrp; ++i) e.vms[m++] = e.vm_pin[(size_t)p->ptoff_host[(size_t)iel] + i];
    }
    e.step = (double)step; e.time = time; e.dt = time_step; e.have_state = true;
  }
  return FSR_OK;
}

// The same for one part: q = [finit; vg] of the step just converged.
int fsr_recovery_update(int base_id, int step, double time, double time_step, const double* q)
{
  return fsr_recovery_update_parts(1, &base_id, step, time, time_step, &q);
}

// partStateVectorSize (solverModule.f90:2183-2206): 3*nnod + 4, -1 for an unknown part, -999 before any part exists
int getPartDeformationStateSize(int bid)
{
  if (g_parts.empty()) return -999;
  auto it = g_parts.find(bid);
  return it == g_parts.end() ? -1 : 3 * it->second.part->nnod + 4;
}

int getPartStressStateSize(int bid)
{
  if (g_parts.empty()) return -999;
  auto it = g_parts.find(bid);
  if (it == g_parts.end()) return -1;
  const int n = (int)it->second.vms.size();
  return n > 0 ? n + 4 : n;
}

static bool save_part(int iop, int bid, double* data, int ndat)
{
  auto it = g_parts.find(bid);
  if (!data || ndat < 4) { set_error("savePartState: state array too small"); return false; }
  // savePartState (solverModule.f90:2230-2263): header, then getDeformation / getStress from position 5
  data[0] = it == g_parts.end() ? 0.0 : it->second.step;
  data[1] = it == g_parts.end() ? 0.0 : it->second.time;
  data[2] = it == g_parts.end() ? 0.0 : it->second.dt;
  data[3] = (double)bid;
  if (it == g_parts.end()) return true;   // like the reference: nothing is copied for a part without recovery
  const Entry& e = it->second;
  const fsr_part* p = e.part;
  if (iop == 1) {
    if (4 + 3 * p->nnod > ndat) { set_error("savePartDeformationState: state array too small (%d < %d)", ndat, 4 + 3 * p->nnod); return false; }
    for (int j = 0; j < p->nnod; ++j) {
      const bool real_node = e.minex.empty() || e.minex[(size_t)j] > 0;   // extra nodes of pinned beams carry no output
      const int d0 = p->madof_host[(size_t)j] - 1;
      for (int k = 0; k < 3; ++k) data[4 + 3 * j + k] = real_node ? e.sv[(size_t)d0 + k] : 0.0;
    }
  } else {
    if (4 + (int)e.vms.size() > ndat) { set_error("savePartStressState: state array too small (%d < %d)", ndat, 4 + (int)e.vms.size()); return false; }
    for (size_t k = 0; k < e.vms.size(); ++k) data[4 + k] = e.vms[k];
  }
  return true;
}

bool savePartDeformationState(int bid, double* data, int ndat) { return save_part(1, bid, data, ndat); }
bool savePartStressState(int bid, double* data, int ndat) { return save_part(2, bid, data, ndat); }

}  // extern "C"
