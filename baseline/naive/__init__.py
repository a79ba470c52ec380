"""GPU comparator (test / bench infrastructure, never imported by gsvc_b200): see naive_rast.cu."""
