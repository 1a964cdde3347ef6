// marin_b200_engine.hpp -- C++ adapter: the B200 engine behind PrMers' abstract `engine` class.
//
// `engine_b200` derives from the reference's `engine` (include/marin/engine.h:16-303) and forwards every
// virtual to the C ABI of libmarin_b200.so (include/marin_b200.h), exactly as the reference's own
// `engine_aevum` adapter forwards to its plugin (src/aevum/EngineAevum.cpp:252-493).  Errors become
// std::runtime_error, as the reference's OpenCL layer throws (include/marin/ocl.h:105-113); the bool
// getters/setters return false only on a size mismatch (engine_gpu.h:2136-2164).
//
// Build against the reference tree:   -DMARIN_B200_USE_REFERENCE_ENGINE_H -I<PrMers>/include
//   (then src/modes/*.cpp compile unchanged; see INTEGRATION.md for the two-line factory change).
// Stand-alone (no GMP, e.g. this repo's tests): the block below declares the same interface minus
// the mpz helpers.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>
#include "marin_b200.h"

#ifdef MARIN_B200_USE_REFERENCE_ENGINE_H
#include "marin/engine.h"
#else
typedef uint32_t uint32;
typedef uint64_t uint64;
class engine
{
protected:
	virtual void get(uint64 *const d, const size_t src) const = 0;
	virtual void set(const size_t dst, uint64 *const d) const = 0;

public:
	typedef size_t Reg;
	engine() {}
	virtual ~engine() {}
	virtual bool is_aevum_backend() const { return false; }
	virtual void release_gpu_resources_for_lowmem_handoff() {}
	virtual size_t get_size() const = 0;
	virtual void sync() const {}
	virtual void set(const Reg dst, const uint32 a) const = 0;
	virtual void copy(const Reg dst, const Reg src) const = 0;
	virtual void square_mul(const Reg src, const uint32 a = 1) const = 0;
	virtual void set_multiplicand(const Reg dst, const Reg src) const = 0;
	virtual void set_multiplicand2(const Reg dst, const Reg src) const { set_multiplicand(dst, src); }
	virtual void mul(const Reg dst, const Reg src, const uint32 a = 1) const = 0;
	virtual void sub(const Reg src, const uint32 a) const = 0;
	virtual void add(const Reg dst, const Reg src) const = 0;
	virtual void mul_add(const Reg dst, const Reg mul_src, const Reg add_src, const uint32 a = 1) const { mul(dst, mul_src, a); add(dst, add_src); }
	virtual void sub_reg(const Reg dst, const Reg src) const = 0;
	virtual void addsub(const Reg s, const Reg d, const Reg a, const Reg b) const { copy(s, a); copy(d, a); add(s, b); sub_reg(d, b); }
	virtual void square_mul_copy(const Reg src, const Reg dst_copy, const uint32 a = 1) const { square_mul(src, a); copy(dst_copy, src); }
	virtual void mul_new(const Reg dst, const Reg src, const uint32 a = 1) const { mul(dst, src, a); }
	virtual void mul_copy(const Reg dst, const Reg src, const Reg dst_copy, const uint32 a = 1) const { mul(dst, src, a); copy(dst_copy, dst); }
	virtual void addsub_copy(const Reg s, const Reg d, const Reg sc, const Reg dc, const Reg a, const Reg b) const { addsub(s, d, a, b); copy(sc, s); copy(dc, d); }
	virtual size_t get_register_data_size() const = 0;
	virtual bool get_data(std::vector<char> &data, const Reg src) const = 0;
	virtual bool set_data(const Reg dst, const std::vector<char> &data) const = 0;
	virtual size_t get_checkpoint_size() const = 0;
	virtual bool get_checkpoint(std::vector<char> &data) const = 0;
	virtual bool set_checkpoint(const std::vector<char> &data) const = 0;
	virtual bool is_equal(const Reg lhs, const Reg rhs) const = 0;
	void pow(const Reg dst, const Reg src, const uint64 e) const
	{
		set_multiplicand(src, src);
		set(dst, 1);
		if (e == 0) return;
		int top = 63; while (top > 0 && !((e >> top) & 1)) --top;
		for (int i = top; i >= 0; --i) { square_mul(dst); if ((e >> i) & 1) mul(dst, src); }
	}
	class digit
	{
		std::vector<uint64> _data;
	public:
		digit(engine *const eng, const Reg src) { _data.resize(eng->get_size()); eng->get(_data.data(), src); }
		size_t get_size() const { return _data.size(); }
		uint32 val(const size_t i) const { return uint32(_data[i]); }
		uint8_t width(const size_t i) const { return uint8_t(_data[i] >> 32); }
		uint64 res64() const
		{
			uint64 r64 = 0; uint8_t s = 0;
			for (uint64 d : _data) { r64 += uint64(uint32(d)) << s; s += uint8_t(d >> 32); if (s >= 64) break; }
			return r64;
		}
		bool equal_to(const uint64 a) const
		{
			uint64 r = a;
			for (uint64 d : _data) { const uint8_t w = uint8_t(d >> 32); if ((r & ((uint64(1) << w) - 1)) != uint32(d)) return false; r >>= w; }
			return true;
		}
		bool equal_to_Mp() const
		{
			for (uint64 d : _data) { const uint8_t w = uint8_t(d >> 32); if (uint32(d) != (uint64(1) << w) - 1) return false; }
			return true;
		}
	};
};
#endif

class engine_b200 : public engine
{
	void *_h;
	uint32_t _q = 0;
	static void ck(int ok) { if (!ok) throw std::runtime_error(std::string("marin_b200 error: ") + mb200_last_error()); }

protected:
	void get(uint64 *const d, const size_t src) const override { ck(mb200_get_digits(_h, src, d, mb200_transform_size(_h))); }
	void set(const size_t dst, uint64 *const d) const override { ck(mb200_set_digits(_h, dst, d, mb200_transform_size(_h))); }

public:
	engine_b200(const uint32_t q, const size_t reg_count, const size_t device, const bool verbose)
		: _h(mb200_create(q, reg_count, uint32_t(device), verbose ? 1 : 0)), _q(q)
	{
		if (!_h) throw std::runtime_error(std::string("marin_b200 error: ") + mb200_last_error());
	}
	~engine_b200() override { mb200_destroy(_h); }
	engine_b200(const engine_b200 &) = delete;
	engine_b200 &operator=(const engine_b200 &) = delete;

	size_t get_size() const override { return mb200_transform_size(_h); }
	void sync() const override { ck(mb200_sync(_h)); }
	void set(const Reg dst, const uint32 a) const override { ck(mb200_set_u32(_h, dst, a)); }
	void copy(const Reg dst, const Reg src) const override { ck(mb200_copy(_h, dst, src)); }
	void square_mul(const Reg src, const uint32 a = 1) const override { ck(mb200_square_mul(_h, src, a)); }
	void set_multiplicand(const Reg dst, const Reg src) const override { ck(mb200_set_multiplicand(_h, dst, src)); }
	// extension (no reference counterpart): dst <- multiplicand of (x + y) from the multiplicands a of x and b of y
	void add_multiplicand(const Reg dst, const Reg a, const Reg b) const { ck(mb200_add_multiplicand(_h, dst, a, b)); }
	void mul(const Reg dst, const Reg src, const uint32 a = 1) const override { ck(mb200_mul(_h, dst, src, a)); }
	void sub(const Reg src, const uint32 a) const override { ck(mb200_sub_u32(_h, src, a)); }
	void add(const Reg dst, const Reg src) const override { ck(mb200_add(_h, dst, src)); }
	void sub_reg(const Reg dst, const Reg src) const override { ck(mb200_sub_reg(_h, dst, src)); }
	void mul_add(const Reg dst, const Reg m, const Reg a, const uint32 k = 1) const override { ck(mb200_mul_add(_h, dst, m, a, k)); }
	void addsub(const Reg s, const Reg d, const Reg a, const Reg b) const override { ck(mb200_addsub(_h, s, d, a, b)); }
	void addsub_copy(const Reg s, const Reg d, const Reg sc, const Reg dc, const Reg a, const Reg b) const override { ck(mb200_addsub_copy(_h, s, d, sc, dc, a, b)); }
	void square_mul_copy(const Reg src, const Reg c, const uint32 a = 1) const override { ck(mb200_square_mul_copy(_h, src, c, a)); }
	void mul_copy(const Reg dst, const Reg src, const Reg c, const uint32 a = 1) const override { ck(mb200_mul_copy(_h, dst, src, c, a)); }
	// extension, no reference counterpart: record a fixed op sequence on a single-CTA engine, replay it as one kernel launch
	void graph_begin() const { ck(mb200_graph_begin(_h)); }
	int graph_end() const { int id = -1; ck(mb200_graph_end(_h, &id)); return id; }
	void graph_launch(const int id) const { ck(mb200_graph_launch(_h, id)); }
	bool is_equal(const Reg lhs, const Reg rhs) const override { int out = 0; ck(mb200_equal(_h, lhs, rhs, &out)); return out != 0; }
#ifdef MARIN_B200_USE_REFERENCE_ENGINE_H
	// the reference's default set_mpz splits z into digits WITHOUT reducing it (engine.h:205-232: bits above the exponent
	// are dropped, so 2^q - 1 + 123 becomes 122); reduce first, as the reference's own plugin adapter does
	// (src/aevum/EngineAevum.cpp:406-414), then let the reference's code do the split
	void set_mpz(const Reg dst, const mpz_t &z) const override
	{
		mpz_t m, r;
		mpz_init(m); mpz_init(r);
		mpz_set_ui(m, 1); mpz_mul_2exp(m, m, _q); mpz_sub_ui(m, m, 1);
		mpz_mod(r, z, m);
		engine::set_mpz(dst, r);
		mpz_clear(m); mpz_clear(r);
	}
#endif

	size_t get_register_data_size() const override { return mb200_register_data_size(_h); }
	bool get_data(std::vector<char> &data, const Reg src) const override
	{
		if (data.size() != get_register_data_size()) return false;
		ck(mb200_get_raw(_h, src, data.data(), data.size())); return true;
	}
	bool set_data(const Reg dst, const std::vector<char> &data) const override
	{
		if (data.size() != get_register_data_size()) return false;
		ck(mb200_set_raw(_h, dst, data.data(), data.size())); return true;
	}
	size_t get_checkpoint_size() const override { return mb200_checkpoint_size(_h); }
	bool get_checkpoint(std::vector<char> &data) const override
	{
		if (data.size() != get_checkpoint_size()) return false;
		ck(mb200_get_checkpoint(_h, data.data(), data.size())); return true;
	}
	bool set_checkpoint(const std::vector<char> &data) const override
	{
		if (data.size() != get_checkpoint_size()) return false;
		ck(mb200_set_checkpoint(_h, data.data(), data.size())); return true;
	}
};

// The factory a PrMers build would call from engine::create_gpu (src/marin/gpu.cpp:52): caller deletes.
inline engine *create_gpu_b200(const uint32_t q, const size_t reg_count, const size_t device, const bool verbose)
{
	return new engine_b200(q, reg_count, device, verbose);
}
