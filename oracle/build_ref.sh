#!/bin/bash
# TEST INFRASTRUCTURE (oracle) -- not part of the product path.
#
# Builds the UNMODIFIED reference (LAMMPS 7Feb2024 + src/OBMD) as a serial
# MPI-STUBS binary from the sources where they lie under /root/reference,
# with plain g++ driven by a small Makefile written here (oracle/Makefile.ref).
# The reference's own build systems (cmake, src/Makefile + Make.sh) are NOT run.
# Outputs go only into oracle/_ref/ (git-ignored, NOT gpurun-ignored):
#   oracle/_ref/lmp_serial      the reference binary (with PLUGIN support)
#   oracle/_ref/gen/*.h         style/package include lists + version stubs
# The only "generated code" LAMMPS needs are include lists (style_*.h,
# packages_*.h) and two tiny stubs (lmpinstalledpkgs.h, lmpgitversion.h); they
# are produced below by grepping the reference headers for the *_CLASS guard.
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
SRC=$REF/code/src
PKGS="DPD-BASIC OBMD MOLECULE PLUGIN"
if [ ! -d "$SRC" ]; then
  echo "build_ref: $SRC not present (GPU box?) -- using prebuilt $OUT if any"; exit 0
fi
mkdir -p "$OUT/gen" "$OUT/obj"
export LC_ALL=C

gen_style () { # $1 guard macro, $2 header prefix, $3 style name
  local tmp="$OUT/gen/style_$3.h.tmp" ptmp="$OUT/gen/packages_$3.h.tmp"
  : > "$tmp"; : > "$ptmp"
  for f in $(cd "$SRC" && grep -sl "$1" $2*.h | sort); do
    echo "#include \"$f\"" >> "$tmp"
  done
  for p in $PKGS; do
    for f in $(cd "$SRC/$p" && grep -sl "$1" $2*.h 2>/dev/null | sort); do
      echo "#include \"$f\"" >> "$tmp"
      printf '#undef PACKAGE\n#define PACKAGE "%s"\n#include "%s"\n' "$p" "$f" >> "$ptmp"
    done
  done
  for t in "$tmp" "$ptmp"; do
    if ! cmp -s "$t" "${t%.tmp}" 2>/dev/null; then mv "$t" "${t%.tmp}"; else rm -f "$t"; fi
  done
}
gen_style ANGLE_CLASS        angle_         angle
gen_style ATOM_CLASS         atom_vec_      atom
gen_style BODY_CLASS         body_          body
gen_style BOND_CLASS         bond_          bond
gen_style COMMAND_CLASS      ""             command
gen_style COMPUTE_CLASS      compute_       compute
gen_style DIHEDRAL_CLASS     dihedral_      dihedral
gen_style DUMP_CLASS         dump_          dump
gen_style FIX_CLASS          fix_           fix
gen_style GRAN_SUB_MOD_CLASS gran_sub_mod_  gran_sub_mod
gen_style IMPROPER_CLASS     improper_      improper
gen_style INTEGRATE_CLASS    ""             integrate
gen_style KSPACE_CLASS       ""             kspace
gen_style MINIMIZE_CLASS     min_           minimize
gen_style NBIN_CLASS         nbin_          nbin
gen_style NPAIR_CLASS        npair_         npair
gen_style NSTENCIL_CLASS     nstencil_      nstencil
gen_style NTOPO_CLASS        ntopo_         ntopo
gen_style PAIR_CLASS         pair_          pair
gen_style READER_CLASS       reader_        reader
gen_style REGION_CLASS       region_        region

cat > "$OUT/gen/lmpinstalledpkgs.h.tmp" <<EOF
#ifndef LMP_INSTALLED_PKGS_H
#define LMP_INSTALLED_PKGS_H
const char * LAMMPS_NS::LAMMPS::installed_packages[] = {
"DPD-BASIC", "MOLECULE", "OBMD", "PLUGIN",
 NULL };
#endif
EOF
cat > "$OUT/gen/lmpgitversion.h.tmp" <<EOF
#ifndef LMP_GIT_VERSION_H
#define LMP_GIT_VERSION_H
bool LAMMPS_NS::LAMMPS::has_git_info() { return false; }
const char *LAMMPS_NS::LAMMPS::git_commit() { return "(unknown)"; }
const char *LAMMPS_NS::LAMMPS::git_branch() { return "(unknown)"; }
const char *LAMMPS_NS::LAMMPS::git_descriptor() { return "(unknown)"; }
#endif
EOF
for t in lmpinstalledpkgs.h lmpgitversion.h; do
  if ! cmp -s "$OUT/gen/$t.tmp" "$OUT/gen/$t" 2>/dev/null; then mv "$OUT/gen/$t.tmp" "$OUT/gen/$t"; else rm -f "$OUT/gen/$t.tmp"; fi
done

make -f "$HERE/Makefile.ref" -j"${JOBS:-$(nproc)}" REF="$REF" OUT="$OUT" PKGS="$PKGS" "$@"
