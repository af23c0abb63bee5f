// lj_xml.h -- a small XML reader for the Mitsuba-style scene files lajolla accepts.
// Stands where the reference uses pugixml (parsers/parse_scene.cpp:1602-1607): elements, attributes, comments,
// the <?xml ?> declaration, character / entity references in attribute values.  Text nodes are skipped -- the scene
// format carries everything in attributes.  Header-only, no dependencies.
#pragma once
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace ljhost {

struct XmlNode {
    std::string name;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<std::unique_ptr<XmlNode>> children;

    bool has(const char *key) const {
        for (auto &a : attrs) if (a.first == key) return true;
        return false;
    }
    // value of an attribute, "" if absent (pugixml's attribute().value() convention)
    const std::string &attr(const char *key) const {
        static const std::string empty;
        for (auto &a : attrs) if (a.first == key) return a.second;
        return empty;
    }
    const XmlNode *child(const char *nm) const {
        for (auto &c : children) if (c->name == nm) return c.get();
        return nullptr;
    }
};

class XmlParser {
public:
    explicit XmlParser(std::string text) : s_(std::move(text)) {}

    // the document's root element
    std::unique_ptr<XmlNode> parse() {
        skip_misc();
        if (p_ >= s_.size() || s_[p_] != '<') fail("no root element");
        auto root = element();
        skip_misc();
        return root;
    }

private:
    std::string s_;
    size_t p_ = 0;

    [[noreturn]] void fail(const std::string &what) const {
        size_t line = 1;
        for (size_t i = 0; i < p_ && i < s_.size(); i++) if (s_[i] == '\n') line++;
        throw std::runtime_error("XML parse error at line " + std::to_string(line) + ": " + what);
    }
    bool starts(const char *lit) const { return s_.compare(p_, strlen(lit), lit) == 0; }
    static bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n'; }
    static bool is_name_char(char c) {
        return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_' || c == '-' || c == '.' || c == ':';
    }
    void skip_space() { while (p_ < s_.size() && is_space(s_[p_])) p_++; }
    void skip_until(const char *lit) {
        size_t e = s_.find(lit, p_);
        if (e == std::string::npos) fail(std::string("unterminated construct, expected ") + lit);
        p_ = e + strlen(lit);
    }
    // whitespace, text, comments, processing instructions, doctype between elements
    void skip_misc() {
        for (;;) {
            while (p_ < s_.size() && s_[p_] != '<') p_++;
            if (p_ >= s_.size()) return;
            if (starts("<!--")) skip_until("-->");
            else if (starts("<?")) skip_until("?>");
            else if (starts("<![CDATA[")) skip_until("]]>");
            else if (starts("<!")) skip_until(">");
            else return;
        }
    }
    std::string name() {
        size_t b = p_;
        while (p_ < s_.size() && is_name_char(s_[p_])) p_++;
        if (p_ == b) fail("name expected");
        return s_.substr(b, p_ - b);
    }
    std::string unescape(const std::string &v) const {
        std::string out;
        out.reserve(v.size());
        for (size_t i = 0; i < v.size(); i++) {
            if (v[i] != '&') { out.push_back(v[i]); continue; }
            size_t e = v.find(';', i);
            if (e == std::string::npos) { out.push_back('&'); continue; }
            std::string ent = v.substr(i + 1, e - i - 1);
            if (ent == "amp") out.push_back('&');
            else if (ent == "lt") out.push_back('<');
            else if (ent == "gt") out.push_back('>');
            else if (ent == "quot") out.push_back('"');
            else if (ent == "apos") out.push_back('\'');
            else if (!ent.empty() && ent[0] == '#') {
                long code = ent.size() > 1 && (ent[1] == 'x' || ent[1] == 'X') ? strtol(ent.c_str() + 2, nullptr, 16) : strtol(ent.c_str() + 1, nullptr, 10);
                if (code > 0 && code < 128) out.push_back((char)code);
            } else { out.append(v, i, e - i + 1); }
            i = e;
        }
        return out;
    }
    std::unique_ptr<XmlNode> element() {
        auto node = std::make_unique<XmlNode>();
        p_++;  // '<'
        node->name = name();
        for (;;) {
            skip_space();
            if (p_ >= s_.size()) fail("unterminated tag");
            if (starts("/>")) { p_ += 2; return node; }
            if (s_[p_] == '>') { p_++; break; }
            std::string key = name();
            skip_space();
            if (p_ >= s_.size() || s_[p_] != '=') fail("'=' expected after attribute " + key);
            p_++;
            skip_space();
            if (p_ >= s_.size() || (s_[p_] != '"' && s_[p_] != '\'')) fail("quoted attribute value expected");
            char q = s_[p_++];
            size_t e = s_.find(q, p_);
            if (e == std::string::npos) fail("unterminated attribute value");
            node->attrs.emplace_back(key, unescape(s_.substr(p_, e - p_)));
            p_ = e + 1;
        }
        for (;;) {  // content up to the matching end tag
            skip_misc();
            if (p_ >= s_.size()) fail("missing </" + node->name + ">");
            if (starts("</")) {
                p_ += 2;
                std::string end = name();
                if (end != node->name) fail("mismatched end tag </" + end + "> for <" + node->name + ">");
                skip_space();
                if (p_ >= s_.size() || s_[p_] != '>') fail("'>' expected");
                p_++;
                return node;
            }
            node->children.push_back(element());
        }
    }
};

inline std::unique_ptr<XmlNode> xml_load_file(const std::string &path) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    std::string text;
    char buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, n);
    fclose(f);
    return XmlParser(std::move(text)).parse();
}

}  // namespace ljhost
