// solver_state.cu -- the in-core recovery state fedempy polls during a dynamics run (host bookkeeping around
// the device calls) and the solver's recovery switches -recovery / -partVMStress / -partDeformation / -frs3file.
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
#include <cstdlib>
#include <map>
#include <string>
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
  bool stress = true;          // mod(recovery, 2) == 1: stress recovery is performed for the part
  bool vms_on = true;          // part%vms is allocated (-partVMStress > 1), possibly with zero length
  fsr_rdb* rdb = nullptr;      // the part's frs3 results database (-partDeformation > 0 or odd -partVMStress)
};
std::map<int, Entry> g_parts;

// The solver's recovery switches (solverInterface.C:447-452).  Until fsr_recovery_options is called the registry keeps
// what it always did: stress recovery on, von Mises state array on, nothing on file.  After the call the switches that
// were not named take the solver's own defaults (-recovery 0, -partDeformation 1, -partVMStress 1).
struct Switches {
  int recovery = 1, part_vms = 2, part_def = 0;
  bool dbl = false;
  std::vector<std::string> frs3;   // one file per recovered part, in registration order (writeRecoveryHeaders getFileName)
  std::string model_file;
  int files_used = 0;
} g_sw;

// "<\"a.frs\",\"b.frs\">" or a single name (getFileName, stressRecoveryModule.f90:860-884)
std::vector<std::string> split_file_list(const std::string& v)
{
  std::vector<std::string> out;
  if (v.empty()) return out;
  if (v[0] != '<') { out.push_back(v); return out; }
  std::string cur;
  bool in = false;
  for (char c : v) {
    if (c == '"') { if (in) out.push_back(cur); cur.clear(); in = !in; }
    else if (in) cur += c;
  }
  return out;
}
}  // namespace

extern "C" {

// The recovery switches of the dynamics solver, as on its command line / in its option files:
//   -recovery N          1 = stress, 2 = gages, 3 = both (stressRecoveryModule.f90: mod(recovery,2) == 1 -> stress recovery)
//   -partVMStress N      0 = off, 1 = to the frs file, 2 = through the state array (savePartStressState), 3 = both
//   -partDeformation N   0 = off, 1 = deformational displacements to the frs file, 2 / 3 = total displacements as well
//   -frs3file NAME       one name, or <"a.frs","b.frs"> with one name per recovered part
//   -double / -fco ... are accepted; unknown options are skipped with their value.
// Applies to the parts registered AFTERWARDS (initRecovery reads the switches once, :583).
int fsr_recovery_options(const char* args)
{
  if (!args) { set_error("fsr_recovery_options: null argument"); return FSR_ERR_ARG; }
  Switches sw;
  sw.recovery = 0; sw.part_vms = 1; sw.part_def = 1;   // the solver's defaults
  std::vector<std::string> tok;
  {
    std::string cur;
    bool quoted = false;
    for (const char* c = args;; ++c) {
      if (*c == '"') quoted = !quoted;
      if (*c == 0 || (!quoted && (*c == ' ' || *c == '\t' || *c == '\n'))) {
        if (!cur.empty()) tok.push_back(cur);
        cur.clear();
        if (*c == 0) break;
      } else cur += *c;
    }
  }
  for (size_t i = 0; i < tok.size(); ++i) {
    const std::string& t = tok[i];
    if (t.size() < 2 || t[0] != '-') continue;
    const std::string key = t.substr(1);
    const bool has_val = i + 1 < tok.size() && !(tok[i + 1].size() > 1 && tok[i + 1][0] == '-' && !isdigit((unsigned char)tok[i + 1][1]));
    auto ival = [&](int& dst) {
      if (!has_val) { set_error("fsr_recovery_options: -%s needs a value", key.c_str()); return false; }
      dst = atoi(tok[++i].c_str());
      return true;
    };
    if (key == "recovery") { if (!ival(sw.recovery)) return FSR_ERR_ARG; }
    else if (key == "partVMStress") { if (!ival(sw.part_vms)) return FSR_ERR_ARG; }
    else if (key == "partDeformation") { if (!ival(sw.part_def)) return FSR_ERR_ARG; }
    else if (key == "frs3file") { if (has_val) sw.frs3 = split_file_list(tok[++i]); }
    else if (key == "fco" || key == "fop" || key == "fao") { if (has_val) ++i; }
    else if (key == "modelfile") { if (has_val) sw.model_file = tok[++i]; }
    else if (key == "double" || key == "double2") sw.dbl = true;
    else if (has_val) ++i;
  }
  if (sw.recovery < 0 || sw.recovery > 3 || sw.part_vms < 0 || sw.part_vms > 3 || sw.part_def < 0 || sw.part_def > 3) {
    set_error("fsr_recovery_options: -recovery / -partVMStress / -partDeformation out of range (%d, %d, %d)", sw.recovery, sw.part_vms, sw.part_def);
    return FSR_ERR_ARG;
  }
  g_sw = sw;
  return FSR_OK;
}

// One FE part of the mechanism with recovery switched on (initRecovery, stressRecoveryModule.f90:560-668).  user_id / descr
// name the part in the frs3 header; sup_tr_init [12] (column-major 3x4, sup%supTrInit) is needed for the total
// displacements of -partDeformation 2 / 3 and may be NULL otherwise.
int fsr_recovery_register_part(int base_id, int user_id, const char* descr, fsr_part* part, const int* minex, const double* sup_tr_init)
{
  if (!part || base_id < 1) { set_error("fsr_recovery_register: bad arguments"); return FSR_ERR_ARG; }
  Entry e;
  e.part = part;
  e.stress = (g_sw.recovery & 1) != 0;
  if (minex) e.minex.assign(minex, minex + part->nnod);
  e.sv.assign((size_t)part->ndof, 0.0);
  const int nvms = fsr_vms_size(part);
  e.vms_on = e.stress && g_sw.part_vms > 1;
  if (e.vms_on) e.vms.assign((size_t)(nvms > 0 ? nvms : 0), 0.0);   // allocate(part%vms) only for writeVMS > 1 (:655-658)
  fsr_recovery_unregister(base_id);
  FSR_CUDA(cudaSetDevice(part->device));
  FSR_CUDA(cudaMallocHost(&e.sv_pin, sizeof(double) * (size_t)std::max(part->ndof, 1)));
  FSR_CUDA(cudaMallocHost(&e.vm_pin, sizeof(double) * (size_t)std::max(part->npts, 1)));
  // writeRecoveryHeaders (:771-815): a results database per recovered part when anything goes to file
  const bool to_file = e.stress && (g_sw.part_def > 0 || (g_sw.part_vms & 1));
  if (to_file && !g_sw.frs3.empty()) {
    const int ifrs = g_sw.files_used++;
    if (g_sw.frs3.size() > 1 && ifrs >= (int)g_sw.frs3.size()) {
      cudaFreeHost(e.sv_pin); cudaFreeHost(e.vm_pin);
      set_error("fsr_recovery_register: too few frs-file names specified (-frs3file)");
      return FSR_ERR_ARG;
    }
    std::string name = g_sw.frs3.size() > 1 ? g_sw.frs3[(size_t)ifrs] : g_sw.frs3[0];
    if (g_sw.frs3.size() == 1 && ifrs > 0) {   // one name for several parts: number them (the solver wants a list; be kind)
      const size_t dot = name.rfind('.');
      name.insert(dot == std::string::npos ? name.size() : dot, "_p" + std::to_string(ifrs + 1));
    }
    fsr_rdb_options o;
    memset(&o, 0, sizeof(o));
    o.out_mask = (g_sw.part_def > 0 ? FSR_OUT_DEFORMATION : 0u) | ((g_sw.part_vms & 1) ? FSR_OUT_VMSTRESS : 0u);
    o.double_precision = g_sw.dbl ? 1 : 0;
    o.part_base_id = base_id;
    o.part_user_id = user_id;
    o.part_descr = descr;
    o.model_file = g_sw.model_file.empty() ? nullptr : g_sw.model_file.c_str();
    o.module_name = "fedem_solver";
    o.minex = minex;
    o.sup_tr_init = g_sw.part_def > 1 ? sup_tr_init : nullptr;
    const int rc = fsr_rdb_create(&e.rdb, part, name.c_str(), &o);
    if (rc < 0) { cudaFreeHost(e.sv_pin); cudaFreeHost(e.vm_pin); return rc; }
  }
  g_parts[base_id] = e;
  return FSR_OK;
}

int fsr_recovery_register(int base_id, fsr_part* part, const int* minex)
{
  return fsr_recovery_register_part(base_id, base_id, nullptr, part, minex, nullptr);
}

int fsr_recovery_unregister(int base_id)
{
  auto it = g_parts.find(base_id);
  if (it == g_parts.end()) return FSR_ERR_ARG;
  int rc = FSR_OK;
  if (it->second.rdb) rc = fsr_rdb_close(it->second.rdb);
  if (it->second.sv_pin) cudaFreeHost(it->second.sv_pin);
  if (it->second.vm_pin) cudaFreeHost(it->second.vm_pin);
  g_parts.erase(it);
  return rc;
}

// closeRecovery (stressRecoveryModule.f90:902-960): the results databases are flushed and closed, the registry is emptied
// and the switches go back to the library's defaults.
int fsr_recovery_close(void)
{
  int rc = FSR_OK;
  while (!g_parts.empty()) {
    const int r = fsr_recovery_unregister(g_parts.begin()->first);
    if (r < 0) rc = r;
  }
  g_sw = Switches();
  return rc;
}

// the frs3 file of a registered part ("" when nothing goes to file); returns the length of the name
int fsr_recovery_file(int base_id, char* buf, int cap)
{
  auto it = g_parts.find(base_id);
  if (it == g_parts.end()) { set_error("fsr_recovery_file: unknown part %d", base_id); return FSR_ERR_ARG; }
  if (!it->second.rdb) { if (buf && cap > 0) buf[0] = 0; return 0; }
  return fsr_rdb_path(it->second.rdb, buf, cap);
}

// The solver's recovery of one converged step for SEVERAL parts (the loop over the parts of
// stressRecoveryModule.f90:1021-1061): the expansion and the stress kernels of all parts are queued on their own streams /
// devices first and waited for afterwards, so the parts of a mechanism overlap instead of running one after the other.
// q[k] = [finit; vg] of part base_ids[k].  The parts' running von Mises envelopes take the step along.
// sup_tr[k] [12]: the part's position at the step (total displacements of -partDeformation 2 / 3), may be NULL;
// do_save = the solver's doSave (results are written at this step): 0 = recover without saving (recoverNotSave).
int fsr_recovery_update_parts_save(int nparts, const int* base_ids, int step, double time, double time_step, const double* const* q,
                                   const double* const* sup_tr, int do_save)
{
  if (nparts < 0 || (nparts > 0 && (!base_ids || !q))) { set_error("fsr_recovery_update_parts: bad arguments"); return FSR_ERR_ARG; }
  std::vector<Entry*> es((size_t)nparts);
  for (int k = 0; k < nparts; ++k) {
    auto it = g_parts.find(base_ids[k]);
    if (it == g_parts.end() || !q[k]) { set_error("fsr_recovery_update: unknown part %d", base_ids[k]); return FSR_ERR_ARG; }
    es[(size_t)k] = &it->second;
  }
  for (int k = 0; k < nparts; ++k) {
    Entry& e = *es[(size_t)k];
    if (!e.stress) continue;   // mod(recovery, 2) == 0: no stress recovery for the part (stressRecovery :1025)
    // without the state array and without von Mises on file only the expansion is needed (recoverNotSave / recoverAndSave)
    const bool want_vm = e.vms_on;
    const int rc = step_enqueue(e.part, q[k], e.sv_pin, want_vm ? e.vm_pin : nullptr);
    if (rc < 0) return rc;
  }
  for (int k = 0; k < nparts; ++k) {
    Entry& e = *es[(size_t)k];
    if (!e.stress) continue;
    const fsr_part* p = e.part;
    int rc = fsr_synchronize(e.part);
    if (rc < 0) return rc;
    std::copy(e.sv_pin, e.sv_pin + p->ndof, e.sv.begin());
    // the in-core vms layout (stressRoutines.f90:324-331): [iel, nenod, nstrp, vm(1..nstrp)] per element with stress points
    size_t m = 0;
    if (e.vms_on)
      for (int iel = 0; iel < p->nel; ++iel) {
        const int nstrp = p->ptoff_host[(size_t)iel + 1] - p->ptoff_host[(size_t)iel];
        if (nstrp <= 0) continue;
        e.vms[m++] = (double)(iel + 1);
        e.vms[m++] = (double)p->nenod_host[(size_t)iel];
        e.vms[m++] = (double)nstrp;
        for (int i = 0; i < nstrp; ++i) e.vms[m++] = e.vm_pin[(size_t)p->ptoff_host[(size_t)iel] + i];
      }
    e.step = (double)step; e.time = time; e.dt = time_step; e.have_state = true;
    if (e.rdb && do_save) {   // writeTimeStepDB + writeDisplacementDB + calcStresses(rdb) of recoverAndSave (:1176-1205)
      if ((rc = fsr_rdb_write_steps(e.rdb, q[k], p->ndim, 1, &step, &time, sup_tr ? sup_tr[k] : nullptr)) < 0) return rc;
    }
  }
  return FSR_OK;
}

int fsr_recovery_update_parts(int nparts, const int* base_ids, int step, double time, double time_step, const double* const* q)
{
  return fsr_recovery_update_parts_save(nparts, base_ids, step, time, time_step, q, nullptr, 1);
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
  if (!it->second.vms_on) return -1;       // part%vms not associated (-partVMStress < 2): getStressSize falls through to -1
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
