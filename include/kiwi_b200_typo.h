/* kiwi_b200 typo image: a flat, pointer-free snapshot of a reference-PREPARED typo transformer
 * (kiwi::PreparedTypoTransformer, /root/reference/include/kiwi/TypoTransformer.h:160-258): the Aho-Corasick pattern
 * trie, the replacement table and the string pool.  Produced by oracle/ref_build/tools/typo_tool.cpp; read by the oracle
 * restatement of generateGraph (oracle/restate/typo.hpp) today and by the lattice kernel when the typo lattice
 * (BASELINE.json config 4) reaches the GPU.  Separate from the model image on purpose: a typo
 * transformer is an analysis option (kiwi_analyze_option_t::typo_transformer), not part of the model. */
#ifndef KIWI_B200_TYPO_H
#define KIWI_B200_TYPO_H
#include <stdint.h>

#define KB2_TYPO_MAGIC 0x314F5059544B42ull   /* "BKTYPO1" */

typedef struct kb2_typo_node {      /* utils::FrozenTrie node, keys ascending per node */
	uint32_t next_offset;
	int32_t  fail;                   /* relative index of the fail node, 0 = none */
	int32_t  value;                  /* index into pats[], -1 = none, -2 = has a sub-match */
	uint16_t num_nexts;
	uint16_t depth;
} kb2_typo_node;

typedef struct kb2_typo_pat { uint32_t repl_off, size, pat_len; } kb2_typo_pat;          /* PatInfo */
typedef struct kb2_typo_repl { uint32_t str_off, length; float cost; uint8_t left_cond; uint8_t pad; uint16_t dialect; } kb2_typo_repl;   /* ReplInfo */

typedef struct kb2_typo_header {
	uint64_t magic;
	uint32_t n_nodes, n_edges, n_pats, n_repls, n_pool;
	float    continual_typo_threshold, lengthening_typo_threshold;   /* INFINITY = disabled */
	uint32_t pad;
	/* followed, each 16-byte aligned and in this order, by:
	 * kb2_typo_node[n_nodes], uint16_t keys[n_edges], int32_t diffs[n_edges], kb2_typo_pat[n_pats], kb2_typo_repl[n_repls], uint16_t pool[n_pool] */
} kb2_typo_header;
#endif
