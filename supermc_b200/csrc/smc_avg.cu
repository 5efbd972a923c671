// smc_avg.cu -- averaged profiles (operation 3): entry points of include/supermc_b200.h.
#include "../../include/supermc_b200.h"
extern "C" int smc_avg_begin(smc_ctx*, int, int, int, int) { return SMC_ERR_STATE; }
extern "C" int smc_avg_run(smc_ctx*, uint64_t, int, smc_event_out*) { return SMC_ERR_STATE; }
extern "C" int smc_avg_run_from_positions(smc_ctx*, int, const smc_event_in*, smc_event_out*) { return SMC_ERR_STATE; }
extern "C" int smc_avg_device_buffer(smc_ctx*, void**, int64_t*) { return SMC_ERR_STATE; }
extern "C" int smc_avg_count(smc_ctx*, int64_t*) { return SMC_ERR_STATE; }
extern "C" int smc_avg_set_count(smc_ctx*, int64_t) { return SMC_ERR_STATE; }
extern "C" int smc_avg_get(smc_ctx*, int, int, int, int, double*) { return SMC_ERR_STATE; }
