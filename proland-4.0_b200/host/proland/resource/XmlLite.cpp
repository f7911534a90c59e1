#include "proland/resource/XmlLite.h"

#include <cctype>

namespace proland
{

const char *XmlElement::Attribute(const char *key) const
{
    for (size_t i = 0; i < attributes.size(); ++i) {
        if (attributes[i].first == key) {
            return attributes[i].second.c_str();
        }
    }
    return NULL;
}

namespace
{

class Parser
{
public:
    explicit Parser(const std::string &t) : s(t), p(0), line(1) {}

    XmlElement document()
    {
        skipMisc();
        if (eof()) fail("no root element");
        XmlElement root = element();
        skipMisc();
        if (!eof()) fail("content after the root element");
        return root;
    }

private:
    const std::string &s;
    size_t p;
    int line;

    bool eof() const { return p >= s.size(); }
    char peek() const { return s[p]; }
    bool starts(const char *lit) const { return s.compare(p, strlen_(lit), lit) == 0; }
    static size_t strlen_(const char *c) { size_t n = 0; while (c[n]) ++n; return n; }
    void advance(size_t n = 1)
    {
        for (size_t i = 0; i < n && p < s.size(); ++i, ++p) {
            if (s[p] == '\n') ++line;
        }
    }
    void fail(const std::string &msg) const
    {
        throw XmlError("XML line " + std::to_string(line) + ": " + msg);
    }
    void skipSpace() { while (!eof() && isspace((unsigned char) peek())) advance(); }
    void skipUntil(const char *end)
    {
        const size_t e = s.find(end, p);
        if (e == std::string::npos) fail(std::string("missing ") + end);
        advance(e + strlen_(end) - p);
    }
    /* white space, comments, processing instructions, DOCTYPE */
    void skipMisc()
    {
        for (;;) {
            skipSpace();
            if (eof()) return;
            if (starts("<!--")) skipUntil("-->");
            else if (starts("<?")) skipUntil("?>");
            else if (starts("<!")) skipUntil(">");
            else return;
        }
    }
    std::string name()
    {
        const size_t b = p;
        while (!eof() && (isalnum((unsigned char) peek()) || peek() == '_' || peek() == '-' || peek() == ':' || peek() == '.')) advance();
        if (p == b) fail("name expected");
        return s.substr(b, p - b);
    }
    static std::string unescape(const std::string &v)
    {
        std::string o;
        for (size_t i = 0; i < v.size(); ++i) {
            if (v[i] != '&') { o += v[i]; continue; }
            static const struct { const char *e; char c; } ents[] = { { "&lt;", '<' }, { "&gt;", '>' }, { "&amp;", '&' }, { "&quot;", '"' }, { "&apos;", '\'' } };
            bool hit = false;
            for (size_t k = 0; k < 5 && !hit; ++k) {
                const size_t n = strlen_(ents[k].e);
                if (v.compare(i, n, ents[k].e) == 0) { o += ents[k].c; i += n - 1; hit = true; }
            }
            if (!hit) o += '&';
        }
        return o;
    }
    XmlElement element()
    {
        XmlElement e;
        e.line = line;
        if (peek() != '<') fail("'<' expected");
        advance();
        e.name = name();
        for (;;) {
            skipSpace();
            if (eof()) fail("unterminated element <" + e.name + ">");
            if (starts("/>")) { advance(2); return e; }
            if (peek() == '>') { advance(); break; }
            const std::string key = name();
            skipSpace();
            if (eof() || peek() != '=') fail("'=' expected after attribute " + key);
            advance();
            skipSpace();
            if (eof() || (peek() != '"' && peek() != '\'')) fail("quoted value expected for attribute " + key);
            const char q = peek();
            advance();
            const size_t b = p;
            while (!eof() && peek() != q) advance();
            if (eof()) fail("unterminated value of attribute " + key);
            e.attributes.push_back(std::make_pair(key, unescape(s.substr(b, p - b))));
            advance();
        }
        /* content: children, comments, text (ignored) until </name> */
        for (;;) {
            if (eof()) fail("missing </" + e.name + ">");
            if (starts("</")) {
                advance(2);
                const std::string close = name();
                if (close != e.name) fail("</" + close + "> closes <" + e.name + ">");
                skipSpace();
                if (eof() || peek() != '>') fail("'>' expected");
                advance();
                return e;
            }
            if (starts("<!--")) { skipUntil("-->"); continue; }
            if (starts("<![CDATA[")) { skipUntil("]]>"); continue; }
            if (starts("<?")) { skipUntil("?>"); continue; }
            if (peek() == '<') { e.children.push_back(element()); continue; }
            advance();
        }
    }
};

}  // namespace

XmlElement parseXml(const std::string &text)
{
    Parser parser(text);
    return parser.document();
}

}  // namespace proland
