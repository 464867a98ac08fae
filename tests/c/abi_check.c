/* Pure C translation unit over include/fi_epp.h: proof that the boundary is C (not merely ctypes-compatible).
 * Built and run by tests/test_boundary.py with `gcc -std=c99 -pedantic -Wall -Wextra -Werror`.
 * Exit code 0: the header compiled as C, the library linked, fi_epp_config_default filled the struct, the
 * YAML loader accepted the reference's prefix-cache document, and fi_epp_create either worked (GPU box) or
 * failed loudly with FI_ERR_CUDA (no device: there is no CPU fallback). */
#include <stdio.h>
#include <string.h>

#include "fi_epp.h"

static const char* kYaml =
    "apiVersion: inference.networking.x-k8s.io/v1alpha1\n"
    "kind: EndpointPickerConfig\n"
    "plugins:\n"
    "- type: prefix-cache-scorer\n"
    "  parameters:\n"
    "    blockSize: 5\n"
    "    maxPrefixBlocksToMatch: 256\n"
    "    lruCapacityPerServer: 31250\n"
    "- type: max-score-picker\n"
    "schedulingProfiles:\n"
    "- name: default\n"
    "  plugins:\n"
    "  - pluginRef: max-score-picker\n"
    "  - pluginRef: prefix-cache-scorer\n"
    "    weight: 100\n";

int main(void) {
  fi_epp_config cfg;
  fi_epp* h = NULL;
  char err[256];
  uint64_t seed = 0;
  int rc;
  if (fi_epp_abi_version() != FI_EPP_ABI_VERSION) return 10;
  if (fi_epp_config_default(&cfg) != FI_OK) return 11;
  if (cfg.struct_size != sizeof(fi_epp_config) || cfg.max_blocks != 256u || cfg.lru_capacity != 31250u) return 12;
  if (fi_epp_config_from_yaml(kYaml, strlen(kYaml), &cfg, err, sizeof err) != FI_OK) {
    fprintf(stderr, "yaml: %s\n", err);
    return 13;
  }
  if (cfg.block_bytes != 5u || cfg.n_profiles != 1u || cfg.profiles[0].scorers[0].weight != 100) return 14;
  if (fi_epp_model_seed("synthetic/model", 15, NULL, 0, &seed) != FI_OK || seed == 0) return 15;
  cfg.num_endpoints = 8;
  cfg.endpoint_count = 8;
  cfg.lru_capacity = 16;
  cfg.max_batch = 4;
  rc = fi_epp_create(&cfg, &h);
  if (rc == FI_OK) {
    fi_epp_destroy(h);
    printf("abi_check: created and destroyed a handle\n");
    return 0;
  }
  if (rc == FI_ERR_CUDA && h == NULL) {
    printf("abi_check: no CUDA device, fi_epp_create failed loudly (%s)\n", fi_epp_status_string(rc));
    return 0;
  }
  return 16;
}
