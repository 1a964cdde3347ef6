// ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
// Exposes the reference's OWN header-only arithmetic (include/marin/arith.h, include/marin/ibdwt.h),
// compiled in place from /root/reference, through a C ABI so that tests can pin the oracle's
// restatement (oracle/marin_oracle.c) against it.  No reference source is copied: this file only
// #includes the headers where they lie.  Output goes to oracle/_ref/ (git-ignored).
// The rest of the reference path (engine.h needs <gmp.h>, engine_gpu.h needs OpenCL + a device,
// kernels/marin.cl needs an OpenCL compiler) is unbuildable in this image; see DESIGN.md.
#include "ibdwt.h"   // pulls arith.h
#include <vector>

extern "C" {

uint64_t ref_mod_add(uint64_t a, uint64_t b) { return mod_add(a, b); }
uint64_t ref_mod_sub(uint64_t a, uint64_t b) { return mod_sub(a, b); }
uint64_t ref_mod_mul(uint64_t a, uint64_t b) { return mod_mul(a, b); }
uint64_t ref_mod_muli(uint64_t a) { return mod_muli(a); }
uint64_t ref_mod_half(uint64_t a) { return mod_half(a); }
uint64_t ref_mod_pow(uint64_t a, uint64_t e) { return mod_pow(a, e); }
uint64_t ref_mod_invert(uint64_t a) { return mod_invert(a); }
uint64_t ref_mod_root_nth(uint64_t n) { return mod_root_nth(n); }

uint32_t ref_adc(uint64_t lhs, uint8_t width, uint64_t *carry) { return adc(lhs, width, *carry); }
uint32_t ref_adc_mul(uint64_t lhs, uint32_t a, uint8_t width, uint64_t *carry) { return adc_mul(lhs, a, width, *carry); }
uint64_t ref_sbc(uint64_t lhs, uint8_t width, uint32_t *carry) { return sbc(lhs, width, *carry); }

size_t ref_transform_size(uint32_t p) { return ibdwt::transform_size(p); }

// weight table in the reference's own layout: pairs (w, 1/w) at index (j % 4) * (n / 4) + j / 4
int ref_weights_widths(size_t n, uint32_t p, uint64_t *weight2n, uint8_t *width)
{
	try { ibdwt::weights_widths(n, p, weight2n, width); } catch (...) { return -1; }
	return 0;
}

// root table (3 n entries), reference layout
void ref_roots(size_t n, uint64_t *root3n) { ibdwt::roots(n, root3n); }

// Host-side canonicalisation of a raw weighted register exactly as engine_gpu_flat::get does it
// (engine_gpu.h:1534-1561), using the reference's adc/mod_mul and its weight layout.
void ref_get_canonical(size_t n, const uint64_t *weight2n, const uint8_t *width, const uint64_t *raw, uint64_t *d)
{
	uint64 c = 0;
	for (size_t k = 0; k < n; ++k)
	{
		const uint64 wi = weight2n[2 * (k / 4 + (k % 4) * (n / 4)) + 1];
		d[k] = adc(mod_mul(raw[k], wi), width[k], c);
	}
	while (c != 0)
	{
		for (size_t k = 0; k < n; ++k)
		{
			d[k] = adc(d[k], width[k], c);
			if (c == 0) break;
		}
	}
	for (size_t k = 0; k < n; ++k) d[k] = uint32(d[k]) | (uint64(width[k]) << 32);
}

}
