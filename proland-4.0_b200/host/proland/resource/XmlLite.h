/*
 * XmlLite -- just enough XML for Proland's resource archives: elements,
 * attributes, comments, the <?xml?> declaration, character data ignored.
 * (The reference reads them with TinyXML through Ork's resource loader, neither
 * of which is in the reference tree.)
 */
#ifndef PROLAND_B200_XML_LITE_H
#define PROLAND_B200_XML_LITE_H

#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace proland
{

struct XmlElement
{
    std::string name;
    std::vector<std::pair<std::string, std::string> > attributes;   /* in document order */
    std::vector<XmlElement> children;
    int line;

    XmlElement() : line(0) {}
    const char *Attribute(const char *key) const;     /* NULL when absent (TinyXML's spelling) */
    const std::string &ValueStr() const { return name; }
};

class XmlError : public std::runtime_error
{
public:
    XmlError(const std::string &what) : std::runtime_error(what) {}
};

/* parses a document; the result is its root element */
XmlElement parseXml(const std::string &text);

}  // namespace proland

#endif
