#include "yaml_lite.h"

#include <cstring>

namespace fwyaml {

const Node* Node::get(std::string_view k) const {
    if (kind != MAP) return nullptr;
    // serde writes the fields in declaration order and the loader asks for them in that order: look first at the entry
    // after the previous answer.  (A mapping with a repeated key — which serde rejects — answers with one of them.)
    const size_t n = children.size();
    if (hint < n && children[hint].key == k) return &children[hint++];
    for (size_t i = 0; i < n; ++i)
        if (children[i].key == k) { hint = (unsigned)i + 1; return &children[i]; }
    return nullptr;
}

namespace {

struct Line {
    int indent;
    const char* s;
    size_t n;
    int no;
};

struct Parser {
    std::vector<Line> lines;
    size_t pos = 0;
    std::string err;
    std::deque<std::string>* pool = nullptr;   // created on first use, handed to the root
    std::string_view keep(std::string&& s) {
        if (!pool) pool = new std::deque<std::string>();
        pool->push_back(std::move(s));
        return pool->back();
    }
    ~Parser() { delete pool; }

    bool fail(int line, const std::string& msg) {
        if (err.empty()) err = "line " + std::to_string(line) + ": " + msg;
        return false;
    }

    // ---- scalars & flow collections (single line) -------------------------------------------------
    static void skip_ws(const char*& p, const char* e) {
        while (p < e && (*p == ' ' || *p == '\t')) ++p;
    }
    // Quoted scalar at p: a view of the text between the quotes when nothing needs unescaping, else of a pooled copy.
    bool parse_quoted(const char*& p, const char* e, std::string_view& view, int line) {
        const char q = *p;
        const char* b = p + 1;
        const char* c = b;
        while (c < e && *c != q && !(q == '"' && *c == '\\')) ++c;
        if (c < e && *c == q && !(q == '\'' && c + 1 < e && c[1] == '\'')) {
            view = std::string_view(b, (size_t)(c - b));
            p = c + 1;
            return true;
        }
        std::string out;
        if (!unescape_quoted(p, e, out, line)) return false;
        view = keep(std::move(out));
        return true;
    }
    bool unescape_quoted(const char*& p, const char* e, std::string& out, int line) {
        char q = *p++;
        out.clear();
        while (p < e) {
            char c = *p++;
            if (c == q) {
                if (q == '\'' && p < e && *p == '\'') { out.push_back('\''); ++p; continue; }
                return true;
            }
            if (q == '"' && c == '\\' && p < e) {
                char d = *p++;
                switch (d) {
                    case 'n': out.push_back('\n'); break;
                    case 't': out.push_back('\t'); break;
                    case '0': out.push_back('\0'); break;
                    default: out.push_back(d); break;
                }
                continue;
            }
            out.push_back(c);
        }
        return fail(line, "unterminated quoted scalar");
    }
    static void set_plain(Node& n, const char* b, const char* e) {
        while (e > b && (e[-1] == ' ' || e[-1] == '\t')) --e;
        std::string_view s(b, (size_t)(e - b));
        if (s.empty() || s == "~" || s == "null" || s == "Null" || s == "NULL") {
            n.kind = Node::NUL;
        } else {
            n.kind = Node::SCALAR;
            n.scalar = s;
        }
    }
    // value inside a flow collection or as a whole line; `stops` = extra terminators in flow context
    // Flow collections nest by recursion; a hostile document ("[[[[...") must not be able to exhaust the stack.
    static constexpr int kMaxFlowDepth = 64;
    bool parse_flow_value(const char*& p, const char* e, Node& n, int line, bool in_flow, int depth = 0) {
        if (depth > kMaxFlowDepth) return fail(line, "flow collections nested too deeply");
        skip_ws(p, e);
        n.line = line;
        if (p >= e) { n.kind = Node::NUL; return true; }
        if (*p == '[') {
            ++p;
            n.kind = Node::SEQ;
            skip_ws(p, e);
            if (p < e && *p == ']') { ++p; return true; }
            for (;;) {
                Node item;
                if (!parse_flow_value(p, e, item, line, true, depth + 1)) return false;
                n.children.push_back(std::move(item));
                skip_ws(p, e);
                if (p < e && *p == ',') { ++p; continue; }
                if (p < e && *p == ']') { ++p; return true; }
                return fail(line, "expected ',' or ']' in flow sequence");
            }
        }
        if (*p == '{') {
            ++p;
            n.kind = Node::MAP;
            skip_ws(p, e);
            if (p < e && *p == '}') { ++p; return true; }
            for (;;) {
                skip_ws(p, e);
                std::string_view key;
                if (p < e && (*p == '"' || *p == '\'')) {
                    if (!parse_quoted(p, e, key, line)) return false;
                } else {
                    const char* b = p;
                    while (p < e && *p != ':' && *p != ',' && *p != '}') ++p;
                    const char* ke = p;
                    while (ke > b && ke[-1] == ' ') --ke;
                    key = std::string_view(b, (size_t)(ke - b));
                }
                skip_ws(p, e);
                if (p >= e || *p != ':') return fail(line, "expected ':' in flow mapping");
                ++p;
                Node val;
                if (!parse_flow_value(p, e, val, line, true, depth + 1)) return false;
                val.key = key;
                n.children.push_back(std::move(val));
                skip_ws(p, e);
                if (p < e && *p == ',') { ++p; continue; }
                if (p < e && *p == '}') { ++p; return true; }
                return fail(line, "expected ',' or '}' in flow mapping");
            }
        }
        if (*p == '"' || *p == '\'') {
            n.kind = Node::SCALAR;
            n.quoted = true;
            return parse_quoted(p, e, n.scalar, line);
        }
        const char* b = p;
        if (in_flow) {
            while (p < e && *p != ',' && *p != ']' && *p != '}') ++p;
        } else {
            p = e;
        }
        set_plain(n, b, p);
        return true;
    }
    bool parse_inline(const char* b, const char* e, Node& n, int line) {
        const char* p = b;
        if (!parse_flow_value(p, e, n, line, false)) return false;
        skip_ws(p, e);
        if (p != e) return fail(line, "trailing characters after value");
        return true;
    }

    // ---- block structure -----------------------------------------------------------------------------
    static bool is_seq_item(const Line& l) { return l.n >= 1 && l.s[0] == '-' && (l.n == 1 || l.s[1] == ' '); }
    // Finds the "key:" split of a block-mapping line. Returns false if the line is not a mapping entry.
    static bool split_key(const Line& l, std::string_view& key, const char*& rest_b, const char*& rest_e) {
        const char* p = l.s;
        const char* e = l.s + l.n;
        if (p < e && (*p == '"' || *p == '\'')) {
            char q = *p;
            const char* b = ++p;
            while (p < e && *p != q) ++p;
            if (p >= e) return false;
            key = std::string_view(b, (size_t)(p - b));
            ++p;
            if (p >= e || *p != ':') return false;
        } else {
            if (p < e && (*p == '[' || *p == '{')) return false;
            const char* b = p;
            while (p < e) {
                if (*p == ':' && (p + 1 == e || p[1] == ' ')) break;
                ++p;
            }
            if (p >= e) return false;
            key = std::string_view(b, (size_t)(p - b));
        }
        ++p;  // ':'
        while (p < e && *p == ' ') ++p;
        rest_b = p;
        rest_e = e;
        return true;
    }

    bool parse_node(int min_indent, Node& out) {
        if (pos >= lines.size() || lines[pos].indent < min_indent) {
            out.kind = Node::NUL;
            return true;
        }
        const Line& l = lines[pos];
        out.line = l.no;
        if (is_seq_item(l)) return parse_seq(l.indent, out);
        std::string_view key;
        const char *rb, *re;
        if (split_key(l, key, rb, re)) return parse_map(l.indent, out);
        // single-line scalar / flow collection
        bool ok = parse_inline(l.s, l.s + l.n, out, l.no);
        ++pos;
        return ok;
    }
    bool parse_seq(int indent, Node& out) {
        out.kind = Node::SEQ;
        out.children.reserve(8);
        while (pos < lines.size() && lines[pos].indent == indent && is_seq_item(lines[pos])) {
            Line& l = lines[pos];
            size_t off = 1;
            while (off < l.n && l.s[off] == ' ') ++off;
            out.children.emplace_back();          // built in place: nothing below touches out.children
            Node& item = out.children.back();
            item.line = l.no;
            if (off >= l.n) {
                ++pos;
                if (!parse_node(indent + 1, item)) return false;
            } else {
                l.indent = indent + (int)off;  // re-read the remainder as a line of its own
                l.s += off;
                l.n -= off;
                if (!parse_node(indent + 1, item)) return false;
            }
        }
        if (pos < lines.size() && lines[pos].indent > indent) return fail(lines[pos].no, "bad indentation in sequence");
        return true;
    }
    bool parse_map(int indent, Node& out) {
        out.kind = Node::MAP;
        out.children.reserve(4);
        while (pos < lines.size() && lines[pos].indent == indent) {
            const Line& l = lines[pos];
            if (is_seq_item(l)) break;
            std::string_view key;
            const char *rb, *re;
            if (!split_key(l, key, rb, re)) return fail(l.no, "expected 'key: value'");
            out.children.emplace_back();
            Node& val = out.children.back();
            val.line = l.no;
            val.key = key;
            if (rb < re) {
                if (!parse_inline(rb, re, val, l.no)) return false;
                ++pos;
            } else {
                ++pos;
                if (pos < lines.size() && lines[pos].indent > indent) {
                    if (!parse_node(indent + 1, val)) return false;
                } else if (pos < lines.size() && lines[pos].indent == indent && is_seq_item(lines[pos])) {
                    if (!parse_seq(indent, val)) return false;  // sequence at the key's own indentation
                } else {
                    val.kind = Node::NUL;
                }
            }
        }
        if (pos < lines.size() && lines[pos].indent > indent) return fail(lines[pos].no, "bad indentation in mapping");
        return true;
    }
};

}  // namespace

bool parse(const char* text, size_t len, Node& root, std::string& err) {
    Parser ps;
    const char* p = text;
    const char* end = text + len;
    int no = 0;
    ps.lines.reserve(len / 12 + 16);
    while (p < end) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        const char* le = nl ? nl : end;
        ++no;
        const char* b = p;
        p = nl ? nl + 1 : end;
        if (le > b && le[-1] == '\r') --le;
        int indent = 0;
        while (b < le && *b == ' ') { ++b; ++indent; }
        if (b < le && *b == '\t') {
            err = "line " + std::to_string(no) + ": tab in indentation";
            return false;
        }
        // strip comments: '#' at start or preceded by a space, outside quotes
        const char* c = b;
        char q = 0;
        const char* ce = le;
        if (!memchr(b, '#', (size_t)(le - b))) c = le;   // no '#' on the line (nearly every line): nothing to strip
        for (; c < le; ++c) {
            if (q) {
                if (*c == q) q = 0;
            } else if (*c == '"' || *c == '\'') {
                // only treat as quote opener at token start
                if (c == b || c[-1] == ' ' || c[-1] == '[' || c[-1] == '{' || c[-1] == ',') q = *c;
            } else if (*c == '#' && (c == b || c[-1] == ' ')) {
                ce = c;
                break;
            }
        }
        while (ce > b && (ce[-1] == ' ' || ce[-1] == '\t')) --ce;
        if (ce == b) continue;
        size_t n = (size_t)(ce - b);
        if (indent == 0 && ((n == 3 && !memcmp(b, "---", 3)) || (n == 3 && !memcmp(b, "...", 3)))) continue;
        if (indent == 0 && n > 4 && !memcmp(b, "--- ", 4)) { b += 4; n -= 4; }
        ps.lines.push_back(Line{indent, b, n, no});
    }
    root = Node();
    if (ps.lines.empty()) return true;
    if (!ps.parse_node(0, root)) {
        err = ps.err;
        return false;
    }
    if (ps.pos != ps.lines.size()) {
        err = "line " + std::to_string(ps.lines[ps.pos].no) + ": unexpected content (indentation?)";
        return false;
    }
    root.pool.reset(ps.pool);
    ps.pool = nullptr;
    return true;
}

}  // namespace fwyaml
