/* qldpc_oracle_osd.c — placeholder translation unit for the OSD restatement. */
#include "qldpc_oracle.h"
