/* Umbrella header: `physisc --b200` output starts with `#define PHYSIS_B200`
 * then includes this (cf. the reference's include/physis/physis.h:10-27). */
#ifndef PHYSIS_PHYSIS_H_
#define PHYSIS_PHYSIS_H_
#if defined(PHYSIS_B200) || !defined(PHYSIS_TARGET_SELECTED)
#include "physis/physis_b200.h"
#endif
#endif
