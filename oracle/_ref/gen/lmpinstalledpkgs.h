#ifndef LMP_INSTALLED_PKGS_H
#define LMP_INSTALLED_PKGS_H
const char * LAMMPS_NS::LAMMPS::installed_packages[] = {
"DPD-BASIC", "MOLECULE", "OBMD", "PLUGIN",
 NULL };
#endif
