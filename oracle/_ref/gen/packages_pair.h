#undef PACKAGE
#define PACKAGE "DPD-BASIC"
#include "pair_dpd.h"
#undef PACKAGE
#define PACKAGE "DPD-BASIC"
#include "pair_dpd_ext.h"
#undef PACKAGE
#define PACKAGE "DPD-BASIC"
#include "pair_dpd_ext_tstat.h"
#undef PACKAGE
#define PACKAGE "DPD-BASIC"
#include "pair_dpd_tstat.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "pair_hbond_dreiding_lj.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "pair_hbond_dreiding_morse.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "pair_lj_charmm_coul_charmm.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "pair_lj_charmm_coul_charmm_implicit.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "pair_lj_charmmfsw_coul_charmmfsh.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "pair_lj_cut_tip4p_cut.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "pair_tip4p_cut.h"
