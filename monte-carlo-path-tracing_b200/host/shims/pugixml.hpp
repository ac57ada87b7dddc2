// pugixml.hpp — TEST INFRASTRUCTURE ONLY (oracle build shim, never shipped).
//
// The reference's parser (src/parser/parser.cpp:8) includes <pugixml.hpp>, which is
// not vendored in the reference tree and not installed here.  This is a from-scratch
// minimal DOM with exactly the subset of the pugixml API that parser.cpp uses
// (parser.cpp:117-133, 183-360, 1439-1617):
//   xml_document::load_file, xml_node::{child, children(), children(name), attribute,
//   name, operator bool}, xml_attribute::{value, as_string, as_float, as_int, as_bool,
//   operator bool}.
// Semantics follow pugixml's documented behaviour: null handles are safe, as_float is
// strtod, as_int is strtol (base 10 / 0x hex), as_bool is true iff the first character
// is one of 1 t T y Y, defaults are returned only for null attributes.
#ifndef ORACLE_SHIM_PUGIXML_HPP
#define ORACLE_SHIM_PUGIXML_HPP

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace pugi {

namespace detail {
struct Node {
    std::string name;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<std::unique_ptr<Node>> children;
};

inline std::string DecodeEntities(const std::string &s) {
    std::string out;
    out.reserve(s.size());
    for (size_t i = 0; i < s.size(); ++i) {
        if (s[i] != '&') { out += s[i]; continue; }
        const size_t semi = s.find(';', i);
        if (semi == std::string::npos) { out += s[i]; continue; }
        const std::string ent = s.substr(i + 1, semi - i - 1);
        if (ent == "amp") out += '&';
        else if (ent == "lt") out += '<';
        else if (ent == "gt") out += '>';
        else if (ent == "quot") out += '"';
        else if (ent == "apos") out += '\'';
        else if (!ent.empty() && ent[0] == '#') {
            const long code = (ent.size() > 1 && (ent[1] == 'x' || ent[1] == 'X'))
                                  ? strtol(ent.c_str() + 2, nullptr, 16)
                                  : strtol(ent.c_str() + 1, nullptr, 10);
            if (code < 128) out += static_cast<char>(code);
        } else { out += s.substr(i, semi - i + 1); }
        i = semi;
    }
    return out;
}

class Parser {
public:
    explicit Parser(const std::string &text) : s_(text), p_(0) {}

    bool Parse(Node *root) {
        std::vector<Node *> stack{root};
        while (p_ < s_.size()) {
            const size_t lt = s_.find('<', p_);
            if (lt == std::string::npos) break;
            p_ = lt;
            if (s_.compare(p_, 4, "<!--") == 0) {
                const size_t e = s_.find("-->", p_ + 4);
                if (e == std::string::npos) return false;
                p_ = e + 3;
            } else if (s_.compare(p_, 2, "<?") == 0) {
                const size_t e = s_.find("?>", p_ + 2);
                if (e == std::string::npos) return false;
                p_ = e + 2;
            } else if (s_.compare(p_, 9, "<![CDATA[") == 0) {
                const size_t e = s_.find("]]>", p_ + 9);
                if (e == std::string::npos) return false;
                p_ = e + 3;
            } else if (s_.compare(p_, 2, "<!") == 0) {
                const size_t e = s_.find('>', p_ + 2);
                if (e == std::string::npos) return false;
                p_ = e + 1;
            } else if (s_.compare(p_, 2, "</") == 0) {
                const size_t e = s_.find('>', p_ + 2);
                if (e == std::string::npos || stack.size() < 2) return false;
                stack.pop_back();
                p_ = e + 1;
            } else {
                ++p_;
                std::unique_ptr<Node> node(new Node());
                node->name = ReadName();
                if (node->name.empty()) return false;
                bool self_closing = false;
                for (;;) {
                    SkipSpace();
                    if (p_ >= s_.size()) return false;
                    if (s_[p_] == '/') { self_closing = true; ++p_; continue; }
                    if (s_[p_] == '>') { ++p_; break; }
                    const std::string key = ReadName();
                    if (key.empty()) return false;
                    SkipSpace();
                    if (p_ >= s_.size() || s_[p_] != '=') return false;
                    ++p_;
                    SkipSpace();
                    if (p_ >= s_.size() || (s_[p_] != '"' && s_[p_] != '\'')) return false;
                    const char quote = s_[p_++];
                    const size_t e = s_.find(quote, p_);
                    if (e == std::string::npos) return false;
                    node->attrs.emplace_back(key, DecodeEntities(s_.substr(p_, e - p_)));
                    p_ = e + 1;
                }
                Node *raw = node.get();
                stack.back()->children.push_back(std::move(node));
                if (!self_closing) stack.push_back(raw);
            }
        }
        return stack.size() == 1;
    }

private:
    void SkipSpace() {
        while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\t' || s_[p_] == '\n' || s_[p_] == '\r')) ++p_;
    }
    std::string ReadName() {
        const size_t b = p_;
        while (p_ < s_.size()) {
            const char c = s_[p_];
            if (c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '=' || c == '>' || c == '/') break;
            ++p_;
        }
        return s_.substr(b, p_ - b);
    }
    const std::string &s_;
    size_t p_;
};
} // namespace detail

class xml_attribute {
public:
    xml_attribute() : v_(nullptr) {}
    explicit xml_attribute(const std::string *v) : v_(v) {}
    explicit operator bool() const { return v_ != nullptr; }
    bool operator!() const { return v_ == nullptr; }
    const char *value() const { return v_ ? v_->c_str() : ""; }
    const char *as_string(const char *def = "") const { return v_ ? v_->c_str() : def; }
    float as_float(float def = 0.0f) const { return v_ ? static_cast<float>(strtod(v_->c_str(), nullptr)) : def; }
    double as_double(double def = 0.0) const { return v_ ? strtod(v_->c_str(), nullptr) : def; }
    int as_int(int def = 0) const {
        if (!v_) return def;
        const char *s = v_->c_str();
        const char *t = s;
        while (*t == ' ' || *t == '\t' || *t == '\n' || *t == '\r') ++t;
        const char *u = (*t == '-' || *t == '+') ? t + 1 : t;
        const int base = (u[0] == '0' && (u[1] == 'x' || u[1] == 'X')) ? 16 : 10;
        return static_cast<int>(strtol(s, nullptr, base));
    }
    bool as_bool(bool def = false) const {
        if (!v_) return def;
        const char c = v_->empty() ? '\0' : (*v_)[0];
        return c == '1' || c == 't' || c == 'T' || c == 'y' || c == 'Y';
    }

private:
    const std::string *v_;
};

class xml_node {
public:
    xml_node() : n_(nullptr) {}
    explicit xml_node(const detail::Node *n) : n_(n) {}
    explicit operator bool() const { return n_ != nullptr; }
    bool operator!() const { return n_ == nullptr; }
    const char *name() const { return n_ ? n_->name.c_str() : ""; }

    xml_node child(const char *name) const {
        if (n_)
            for (const auto &c : n_->children)
                if (c->name == name) return xml_node(c.get());
        return xml_node();
    }
    xml_attribute attribute(const char *name) const {
        if (n_)
            for (const auto &a : n_->attrs)
                if (a.first == name) return xml_attribute(&a.second);
        return xml_attribute();
    }
    std::vector<xml_node> children() const {
        std::vector<xml_node> out;
        if (n_)
            for (const auto &c : n_->children) out.emplace_back(c.get());
        return out;
    }
    std::vector<xml_node> children(const char *name) const {
        std::vector<xml_node> out;
        if (n_)
            for (const auto &c : n_->children)
                if (c->name == name) out.emplace_back(c.get());
        return out;
    }

protected:
    const detail::Node *n_;
};

struct xml_parse_result {
    bool ok = false;
    explicit operator bool() const { return ok; }
};

class xml_document : public xml_node {
public:
    xml_document() : root_(new detail::Node()) { n_ = root_.get(); }
    xml_parse_result load_file(const char *path) {
        xml_parse_result r;
        FILE *f = fopen(path, "rb");
        if (!f) return r;
        std::string text;
        char buf[65536];
        size_t n;
        while ((n = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, n);
        fclose(f);
        root_.reset(new detail::Node());
        n_ = root_.get();
        detail::Parser parser(text);
        r.ok = parser.Parse(root_.get());
        return r;
    }

private:
    std::unique_ptr<detail::Node> root_;
};

} // namespace pugi

#endif
