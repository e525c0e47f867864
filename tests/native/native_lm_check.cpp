// CPU check (tests/test_native_lm.py): the native readers of the reference's sj.knlm / skipbigram.mdl files (kiwi_b200/csrc/native_lm.cpp,
// no reference library involved) against the sections of the model images that flatten_model dumped from the reference's own in-memory
// models - byte for byte, including the suffix links, the BOS state and unk_ll that the reference computes at load time.
//   native_lm_check <libkiwi_b200.so> <knlm image> <sj.knlm> [<sbg image> <skipbigram.mdl> [<cong image> <cong.mdl>]]
//   (pass - - for a pair to skip)
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <fstream>
#include <vector>
#include "../../include/kiwi_b200.h"
#include "../../include/kiwi_b200_image.h"

static std::vector<char> readAll(const char* p) { std::ifstream f{ p, std::ios::binary }; return std::vector<char>{ std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>() }; }

static int cmp(const char* what, const std::vector<char>& img, const kb2_header* h, int sec, const char* blob, uint64_t off, uint64_t bytes)
{
	if (h->sec[sec].nbytes != bytes) { std::printf("%s: %llu bytes in the image, %llu native\n", what, (unsigned long long)h->sec[sec].nbytes, (unsigned long long)bytes); return 1; }
	if (bytes && std::memcmp(img.data() + h->sec[sec].offset, blob + off, bytes)) { std::printf("%s: contents differ\n", what); return 1; }
	std::printf("%s: %llu bytes identical\n", what, (unsigned long long)bytes);
	return 0;
}

int main(int argc, char** argv)
{
	if (argc < 4) return 2;
	void* lib = dlopen(argv[1], RTLD_NOW);
	if (!lib) { std::printf("dlopen: %s\n", dlerror()); return 2; }
	auto knlm = (int (*)(const char*, void**, uint64_t*))dlsym(lib, "kiwi_b200_native_knlm");
	auto sbg = (int (*)(const char*, void**, uint64_t*))dlsym(lib, "kiwi_b200_native_sbg");
	auto cong = (int (*)(const char*, void**, uint64_t*))dlsym(lib, "kiwi_b200_native_cong");
	auto err = (const char* (*)())dlsym(lib, "kiwi_b200_native_error");
	auto freeFn = (void (*)(void*))dlsym(lib, "kiwi_b200_free");
	if (!knlm || !sbg || !cong || !err || !freeFn) { std::printf("missing symbols\n"); return 2; }
	int bad = 0;
	{
		const auto img = readAll(argv[2]);
		const kb2_header* h = reinterpret_cast<const kb2_header*>(img.data());
		void* blob = nullptr; uint64_t size = 0;
		if (knlm(argv[3], &blob, &size)) { std::printf("native knlm failed: %s\n", err()); return 1; }
		const auto* n = static_cast<const kiwi_b200_native_knlm_t*>(blob); const char* b = static_cast<const char*>(blob);
		bad += cmp("KN_NODES", img, h, KB2_SEC_KN_NODES, b, n->nodes_off, n->nodes_bytes);
		bad += cmp("KN_KEYS", img, h, KB2_SEC_KN_KEYS, b, n->keys_off, n->keys_bytes);
		bad += cmp("KN_VALUES", img, h, KB2_SEC_KN_VALUES, b, n->values_off, n->values_bytes);
		bad += cmp("KN_ROOT", img, h, KB2_SEC_KN_ROOT, b, n->root_off, n->root_bytes);
		bad += cmp("KN_HTX", img, h, KB2_SEC_KN_HTX, b, n->htx_off, n->htx_bytes);
		const bool sc = n->num_nodes == h->kn_num_nodes && n->num_edges == h->kn_num_edges && n->htx_vocab == h->kn_htx_vocab && n->has_htx == h->kn_has_htx
			&& n->order == h->kn_order && n->vocab_size == h->lang_vocab_size && n->bos_node == h->kn_bos_node && std::memcmp(&n->unk_ll, &h->kn_unk_ll, 4) == 0;
		std::printf("knlm scalars %s (nodes %u edges %u htx vocab %u order %u bos %d unk_ll %a / image %d %a)\n", sc ? "identical" : "DIFFER", n->num_nodes, n->num_edges, n->htx_vocab, n->order, n->bos_node, n->unk_ll, h->kn_bos_node, h->kn_unk_ll);
		bad += sc ? 0 : 1;
		freeFn(blob);
	}
	if (argc >= 6 && std::strcmp(argv[4], "-"))
	{
		const auto img = readAll(argv[4]);
		const kb2_header* h = reinterpret_cast<const kb2_header*>(img.data());
		void* blob = nullptr; uint64_t size = 0;
		if (sbg(argv[5], &blob, &size)) { std::printf("native sbg failed: %s\n", err()); return 1; }
		const auto* n = static_cast<const kiwi_b200_native_sbg_t*>(blob); const char* b = static_cast<const char*>(blob);
		bad += cmp("SB_PTRS", img, h, KB2_SEC_SB_PTRS, b, n->ptrs_off, n->ptrs_bytes);
		bad += cmp("SB_KEYS", img, h, KB2_SEC_SB_KEYS, b, n->keys_off, n->keys_bytes);
		bad += cmp("SB_COMPS", img, h, KB2_SEC_SB_COMPS, b, n->comps_off, n->comps_bytes);
		bad += cmp("SB_DISCNTS", img, h, KB2_SEC_SB_DISCNTS, b, n->discnts_off, n->discnts_bytes);
		bad += cmp("SB_VALID", img, h, KB2_SEC_SB_VALID, b, n->valid_off, n->valid_bytes);
		const bool sc = n->vocab_size == h->sb_vocab_size && n->window_size == h->sb_window_size && n->num_pairs == h->sb_num_pairs;
		std::printf("sbg scalars %s\n", sc ? "identical" : "DIFFER");
		bad += sc ? 0 : 1;
		freeFn(blob);
	}
	if (argc >= 8 && std::strcmp(argv[6], "-"))
	{
		const auto img = readAll(argv[6]);
		const kb2_header* h = reinterpret_cast<const kb2_header*>(img.data());
		void* blob = nullptr; uint64_t size = 0;
		if (cong(argv[7], &blob, &size)) { std::printf("native cong failed: %s\n", err()); return 1; }
		const auto* n = static_cast<const kiwi_b200_native_cong_t*>(blob); const char* b = static_cast<const char*>(blob);
		bad += cmp("CG_NODES", img, h, KB2_SEC_CG_NODES, b, n->nodes_off, n->nodes_bytes);
		bad += cmp("CG_KEYS", img, h, KB2_SEC_CG_KEYS, b, n->keys_off, n->keys_bytes);
		bad += cmp("CG_VALUES", img, h, KB2_SEC_CG_VALUES, b, n->values_off, n->values_bytes);
		bad += cmp("CG_ROOT", img, h, KB2_SEC_CG_ROOT, b, n->root_off, n->root_bytes);
		bad += cmp("CG_CTX_EMB", img, h, KB2_SEC_CG_CTX_EMB, b, n->ctx_emb_off, n->ctx_emb_bytes);
		bad += cmp("CG_OUT_EMB", img, h, KB2_SEC_CG_OUT_EMB, b, n->out_emb_off, n->out_emb_bytes);
		bad += cmp("CG_INV_VOCAB", img, h, KB2_SEC_CG_INV_VOCAB, b, n->inv_vocab_off, n->inv_vocab_bytes);
		bad += cmp("CG_OUT_BIAS", img, h, KB2_SEC_CG_OUT_BIAS, b, n->out_bias_off, n->out_bias_bytes);
		const bool sc = n->num_nodes == h->cg_num_nodes && n->num_edges == h->cg_num_edges && n->root_size == h->cg_root_size && n->dim == h->cg_dim
			&& n->context_size == h->cg_context_size && n->key_size == h->cg_key_size && n->flags == h->cg_flags && n->vocab_size == h->lang_vocab_size;
		std::printf("cong scalars %s (nodes %u edges %u dim %u contexts %u key size %u)\n", sc ? "identical" : "DIFFER", n->num_nodes, n->num_edges, n->dim, n->context_size, n->key_size);
		bad += sc ? 0 : 1;
		freeFn(blob);
	}
	std::printf("mismatching sections %d\n", bad);
	return bad ? 1 : 0;
}
