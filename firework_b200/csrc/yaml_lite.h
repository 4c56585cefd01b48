// yaml_lite — a small parser for the YAML subset serde_yaml 0.8 emits (and hand-edited variants of it):
// block mappings / sequences by indentation, plain / quoted scalars, `~`/null, `[]` `{}` and simple
// one-line flow collections, `#` comments, a leading `---`.  No anchors, tags, multi-docs or block
// scalars: firework scene documents (reference: serde derive on src/scene.rs:19-24, 268-277) use none.
#pragma once
#include <cstddef>
#include <deque>
#include <memory>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

namespace fwyaml {

// Scalars and keys are views: into the source text (which must outlive the tree) for plain and escape-free quoted scalars,
// into the root's `pool` for the rare quoted scalar that needed unescaping.  A scene document is ~10^4 scalars; owning
// strings and per-node key/value vectors made the parse allocation-bound.
struct Node {
    enum Kind : unsigned char { SCALAR, MAP, SEQ, NUL } kind = NUL;
    bool quoted = false;                                  // SCALAR written with quotes (never null/bool)
    int line = 0;                                         // 1-based source line, for error messages
    mutable unsigned hint = 0;                            // get(): where the previous lookup ended
    std::string_view scalar;                              // SCALAR
    std::string_view key;                                 // this node's key when its parent is a MAP
    std::vector<Node> children;                           // MAP entries / SEQ items, in document order
    std::unique_ptr<std::deque<std::string>> pool;        // root only: storage of unescaped scalars

    const Node* get(std::string_view k) const;            // MAP lookup, nullptr if absent / not a map
    bool is_null() const { return kind == NUL; }
    std::string str() const { return std::string(scalar); }
};

// Parses `text`; on failure returns false and fills `err` ("line N: ...").  `root` refers to `text`.
bool parse(const char* text, size_t len, Node& root, std::string& err);

}  // namespace fwyaml
