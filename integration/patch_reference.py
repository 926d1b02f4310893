#!/usr/bin/env python
"""Apply the seam-by-seam patch of INTEGRATION.md to a copy of the reference (fedluc/HSMC) sources.

    python integration/patch_reference.py /path/to/HSMC/src  /tmp/hsmc_patched  [--build]

Nothing of the reference is stored in this repository: the script copies the caller's `src/` tree,
finds each seam function by its NAME, and replaces its body with a call into `libhsmc_gpu.so`
(`include/hsmc_gpu.h`).  The replacement bodies below are this repository's own code; everything else of
the reference (input parser, box/lattice set-up, optimizer, output writers, `exec.c`) is left as it is.
`--build` compiles the result with the author's flags (`gcc -O2 -std=gnu99`), against GSL if present or
the test shim under `oracle/gsl_shim` otherwise, and links it to the CUDA library.

Seams (reference file:line -> ABI call), see INTEGRATION.md for the table:
  cell_list.c:33  cell_list_init      -> hsmc_gpu_create + hsmc_gpu_upload
  cell_list.c:80  cell_list_free      -> hsmc_gpu_destroy
  nvt.c:201       sweep_nvt           -> hsmc_gpu_sweep_nvt
  npt.c:177       sweep_npt           -> hsmc_gpu_sweep_nvt + Binomial(N, 1/(N+1)) volume moves
  moves.c:83      vol_move            -> hsmc_gpu_overlap_scaled / hsmc_gpu_rescale / hsmc_gpu_add_vol_move
  moves.c:229,243 get/reset counters  -> hsmc_gpu_counters / hsmc_gpu_reset_counters
  compute_widom_chem_pot.c:44 widom_insertion   -> hsmc_gpu_widom
  compute_rdf.c:110           rdf_hist_compute  -> hsmc_gpu_rdf_counts
  compute_press.c:123,243     pressv/presst_compute_hist -> hsmc_gpu_contact_counts / hsmc_gpu_presst_flags
  compute_order_parameter.c:84 global_ql_compute -> hsmc_gpu_order_parameter
  io_config.c:28,134          write_restart / write_config -> hsmc_gpu_download first
"""
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

GLUE_DECL = r"""
/* ---- added by integration/patch_reference.py: the GPU handle shared by the patched seams ---- */
#include "hsmc_gpu.h"
hsmc_gpu *gpu_handle(void);
void gpu_check(int rc);
"""

GLUE_DEF = r"""
/* ---- added by integration/patch_reference.py ---- */
static hsmc_gpu *G_GPU = NULL;
hsmc_gpu *gpu_handle(void) { return G_GPU; }
/* the reference's error convention: message on stdout, exit(EXIT_FAILURE) (cell_list.c:127-128) */
void gpu_check(int rc) {
  if (rc) { printf("ERROR: %s\n", hsmc_gpu_last_error()); exit(EXIT_FAILURE); }
}
"""

BODIES = {
    ("cell_list.c", "cell_list_init"): r"""
  /* host-side cell geometry stays available to get_cell_list_info() users */
  compute_cell_list_info(true);
  (void)alloc;
  box_info b = sim_box_info_get();
  double L[3] = {b.lx, b.ly, b.lz};
  /* Same box as the live handle (the NVT restart from the lattice after the optimizer, nvt.c:58-62):
     keep the handle, so the device Philox sweep counter runs on as in the drop-in driver.  A different
     box (first call; NpT restart after the optimizer, npt.c:54-61) needs a fresh handle. */
  if (G_GPU) {
    hsmc_gpu_info gi;
    gpu_check(hsmc_gpu_get_info(G_GPU, &gi));
    if (gi.box[0] != L[0] || gi.box[1] != L[1] || gi.box[2] != L[2] || gi.n_total != part_info_get().NN) {
      hsmc_gpu_destroy(G_GPU);
      G_GPU = NULL;
    }
  }
  if (!G_GPU) {
    hsmc_gpu_config cfg = {0};
    cfg.world = 1;
    cfg.seed = G_IN.seed;
    cfg.cell_min = G_IN.neigh_dr;
    cfg.regrid_interval = 1;
    gpu_check(hsmc_gpu_create(&G_GPU, &cfg, part_info_get().NN, L));
  }
  gpu_check(hsmc_gpu_upload(G_GPU, &part_config_get()[0][0], part_info_get().NN));
""",
    ("cell_list.c", "cell_list_free"): r"""
  if (G_GPU) hsmc_gpu_destroy(G_GPU);
  G_GPU = NULL;
""",
    ("nvt.c", "sweep_nvt"): r"""
  /* N trial displacements = one checkerboard sweep on the device */
  gpu_check(hsmc_gpu_sweep_nvt(gpu_handle(), 1, G_IN.dr_max));
""",
    ("npt.c", "sweep_npt"): r"""
  /* the reference draws, for each of N steps, a volume move with probability 1/(N+1); here the particle
     moves are one device sweep and the number of volume moves of the sweep is drawn from the same
     Binomial(N, 1/(N+1)) law by geometric skipping */
  gpu_check(hsmc_gpu_sweep_nvt(gpu_handle(), 1, G_IN.dr_max));
  int N = part_info_get().NN;
  double log_q = log1p(-1.0 / (N + 1.0));
  long pos = 0;
  for (;;) {
    double u = rng_get_double();
    if (u <= 0.0) break;
    pos += (long)floor(log(u) / log_q) + 1;
    if (pos > N) break;
    vol_move();
  }
""",
    ("moves.c", "vol_move"): r"""
  /* same steps as the CPU version: ln V perturbation, global overlap verdict under the isotropic
     scaling, Metropolis test with the (N+1) ln(V'/V) measure, then box + coordinates + cell list */
  int N = part_info_get().NN;
  box_info box = sim_box_info_get();
  double r_dv = rng_get_double();
  double vol_new = exp(log(box.vol) + (r_dv - 0.5) * G_IN.dv_max);
  double vol_ratio = vol_new / box.vol;
  double sf = pow(vol_ratio, 1. / 3.);
  int overlap = 0, accepted = 0;
  gpu_check(hsmc_gpu_overlap_scaled(gpu_handle(), sf, &overlap));
  if (!overlap) {
    double boltz_fact = exp(G_IN.press * (box.vol - vol_new) + (N + 1) * log(vol_ratio));
    double r_acc = rng_get_double();
    if (r_acc < boltz_fact) {
      accepted = 1;
      G_IN.rho = N / vol_new;
      sim_box_init(G_IN.type, G_IN.nx, G_IN.ny, G_IN.nz, G_IN.rho);
      box = sim_box_info_get();
      double L[3] = {box.lx, box.ly, box.lz};
      gpu_check(hsmc_gpu_rescale(gpu_handle(), sf, L));
    }
  }
  gpu_check(hsmc_gpu_add_vol_move(gpu_handle(), accepted));
""",
    ("moves.c", "get_moves_counters"): r"""
  int64_t c[6];
  gpu_check(hsmc_gpu_counters(gpu_handle(), c));
  if (pm != NULL) *pm = (int)c[0];
  if (apm != NULL) *apm = (int)c[1];
  if (rpm != NULL) *rpm = (int)c[2];
  if (vm != NULL) *vm = (int)c[3];
  if (avm != NULL) *avm = (int)c[4];
  if (rvm != NULL) *rvm = (int)c[5];
""",
    ("moves.c", "reset_moves_counters"): r"""
  if (gpu_handle()) gpu_check(hsmc_gpu_reset_counters(gpu_handle()));
""",
    ("compute_widom_chem_pot.c", "widom_insertion"): r"""
  static uint64_t sample = 0;
  int64_t accepted = 0;
  gpu_check(hsmc_gpu_widom(gpu_handle(), sample++, 0, G_IN.mu_insertions, 1, &accepted));
  wtest = (int)accepted;
  mu = (wtest > 0) ? -log((double)wtest / G_IN.mu_insertions) : 0.0;
""",
    ("compute_rdf.c", "rdf_hist_compute"): r"""
  uint64_t *c = calloc(rdf_hist_nn > 0 ? rdf_hist_nn : 1, sizeof(uint64_t));
  if (rdf_hist_nn > 0) gpu_check(hsmc_gpu_rdf_counts(gpu_handle(), G_IN.rdf_dr, rdf_hist_nn, c));
  for (int k = 0; k < rdf_hist_nn; k++) rdf_hist[k] += 2.0 * (double)c[k];
  free(c);
""",
    ("compute_press.c", "pressv_compute_hist"): r"""
  uint64_t *c = calloc(pressv_hist_nn > 0 ? pressv_hist_nn : 1, sizeof(uint64_t));
  gpu_check(hsmc_gpu_contact_counts(gpu_handle(), G_IN.pressv_dr, pressv_hist_nn, c));
  for (int k = 0; k < pressv_hist_nn; k++) pressv_hist[k] += 2.0 * (double)c[k];
  free(c);
""",
    ("compute_press.c", "presst_compute_hist"): r"""
  if (presst_hist_nn <= 0) return;
  double *sf = malloc(sizeof(double) * presst_hist_nn);
  int *free_of_overlap = calloc(presst_hist_nn, sizeof(int));
  for (int k = 0; k < presst_hist_nn; k++) sf[k] = pow(1 - presst_xi[k], 1. / 3.);
  gpu_check(hsmc_gpu_presst_flags(gpu_handle(), sf, presst_hist_nn, free_of_overlap));
  for (int k = 0; k < presst_hist_nn; k++)
    if (free_of_overlap[k]) presst_hist[k] += 1.0;
  free(sf);
  free(free_of_overlap);
""",
    ("compute_order_parameter.c", "global_ql_compute"): r"""
  /* bonds within ql_rmax; the device's (even-count) cell grid bounds the cutoff like the host's does */
  hsmc_gpu_info gi;
  gpu_check(hsmc_gpu_get_info(gpu_handle(), &gi));
  double edge = fmin(gi.cell_size[0], fmin(gi.cell_size[1], gi.cell_size[2]));
  double rmax = G_IN.ql_rmax < edge ? G_IN.ql_rmax : edge;
  gpu_check(hsmc_gpu_order_parameter(gpu_handle(), G_IN.ql_order, rmax, &ql_ave));
""",
}

# seams that keep their body and only get a statement in front of it
PREPEND = {
    ("io_config.c", "write_restart"): "  gpu_check(hsmc_gpu_download(gpu_handle(), &part_config_get()[0][0]));\n",
    ("io_config.c", "write_config"): "  gpu_check(hsmc_gpu_download(gpu_handle(), &part_config_get()[0][0]));\n",
}


def find_body(text, name):
    """(start, end) of the text between the braces of the top-level definition of `name`"""
    m = re.search(r"^[A-Za-z_][\w \*]*\b%s\s*\([^;{]*\)\s*\{" % re.escape(name), text, re.M)
    if not m:
        raise RuntimeError(f"definition of {name} not found")
    i = m.end()
    depth, j = 1, i
    while depth:
        ch = text[j]
        if ch == "{":
            depth += 1
        elif ch == "}":
            depth -= 1
        elif ch == '"':                      # skip string literals
            j += 1
            while text[j] != '"':
                j += 2 if text[j] == "\\" else 1
        elif text.startswith("//", j):
            j = text.index("\n", j)
        elif text.startswith("/*", j):
            j = text.index("*/", j) + 1
        j += 1
    return i, j - 1


def patch_tree(src, dst):
    if os.path.exists(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst)
    files = {}
    for (fname, func), body in BODIES.items():
        text = files.get(fname) or open(os.path.join(dst, fname)).read()
        a, b = find_body(text, func)
        files[fname] = text[:a] + "\n  /* body replaced by integration/patch_reference.py */" + body + text[b:]
    for (fname, func), stmt in PREPEND.items():
        text = files.get(fname) or open(os.path.join(dst, fname)).read()
        a, _ = find_body(text, func)
        files[fname] = text[:a] + "\n" + stmt + text[a:]
    # the handle lives in cell_list.c; every patched file includes cell_list.h (io_config.c gets it added)
    files["cell_list.c"] = re.sub(r'(#include "cell_list.h"\n)', r"\1" + GLUE_DEF.replace("\\", "\\\\"), files["cell_list.c"], count=1)
    hdr = open(os.path.join(dst, "cell_list.h")).read()
    k = hdr.rindex("#endif")
    files["cell_list.h"] = hdr[:k] + GLUE_DECL + "\n" + hdr[k:]
    for fname in ("io_config.c", "npt.c", "compute_order_parameter.c", "compute_rdf.c", "compute_press.c"):
        text = files.get(fname) or open(os.path.join(dst, fname)).read()
        extra = '#include <math.h>\n#include <stdint.h>\n#include "sim_info.h"\n#include "cell_list.h"\n'
        files[fname] = extra + text
    for fname, text in files.items():
        with open(os.path.join(dst, fname), "w") as f:
            f.write(text)
    return sorted(files)


def build(dst):
    inc = ["-I", os.path.join(ROOT, "include")]
    if not os.path.exists("/usr/include/gsl/gsl_rng.h"):
        inc += ["-I", os.path.join(ROOT, "oracle", "gsl_shim")]        # test shim: MT19937 + sphPlm only
        libs = []
    else:
        libs = ["-lgsl", "-lgslcblas"]
    lib = os.path.join(ROOT, "hsmc_b200", "csrc")
    srcs = sorted(os.path.join(dst, f) for f in os.listdir(dst) if f.endswith(".c"))
    exe = os.path.join(dst, "hsmc_gpu_patched")
    cmd = ["gcc", "-O2", "-std=gnu99", "-w", *inc, "-I", dst, *srcs, "-o", exe, "-L", lib, "-lhsmc_gpu",
           f"-Wl,-rpath,{lib}", *libs, "-lz", "-lm"]
    subprocess.run(cmd, check=True)
    return exe


def main():
    if len(sys.argv) < 3:
        print(__doc__)
        return 2
    changed = patch_tree(sys.argv[1], sys.argv[2])
    print("patched:", ", ".join(changed))
    if "--build" in sys.argv[3:]:
        print("built:", build(sys.argv[2]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
