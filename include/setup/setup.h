/* Forwarding header: the reference keeps this declaration set in src/setup/setup.h; here everything lives in
 * include/ckzg.h (one umbrella header, no blst dependency). */
#include "../ckzg.h"
