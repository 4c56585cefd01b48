// yaml_lite — a small parser for the YAML subset serde_yaml 0.8 emits (and hand-edited variants of it):
// block mappings / sequences by indentation, plain / quoted scalars, `~`/null, `[]` `{}` and simple
// one-line flow collections, `#` comments, a leading `---`.  No anchors, tags, multi-docs or block
// scalars: firework scene documents (reference: serde derive on src/scene.rs:19-24, 268-277) use none.
#pragma once
#include <cstddef>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace fwyaml {

struct Node {
    enum Kind { SCALAR, MAP, SEQ, NUL } kind = NUL;
    std::string scalar;                                   // SCALAR
    bool quoted = false;                                  // SCALAR written with quotes (never null/bool)
    std::vector<std::pair<std::string, Node>> map;        // MAP (ordered)
    std::vector<Node> seq;                                // SEQ
    int line = 0;                                         // 1-based source line, for error messages

    const Node* get(const char* key) const;               // MAP lookup, nullptr if absent / not a map
    bool is_null() const { return kind == NUL; }
};

// Parses `text`; on failure returns false and fills `err` ("line N: ...").
bool parse(const char* text, size_t len, Node& root, std::string& err);

}  // namespace fwyaml
