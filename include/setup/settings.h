/* Forwarding header: the reference keeps this declaration set in src/setup/settings.h; here everything lives in
 * include/ckzg.h (one umbrella header, no blst dependency). */
#include "../ckzg.h"
