// ref_main.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles the UNMODIFIED reference translation unit
// (/root/reference/src/smoothMesh.C, read where it lies; never copied) against the OpenFOAM facade in
// this directory.  Output: oracle/_ref/smoothMesh_ref, an executable with the reference's own main(),
// command line and log lines.  See OpenFOAMFacade.H for what the facade does and does not prove, and
// FacadeSupport.H for why nothing else of this repository is included here.
//
// Build recipe: oracle/Makefile.ref (only in the build container, where /root/reference exists).
#include <dirent.h>

#include "FacadeMesh.H"

namespace Foam
{
// "key value;" entries of a dictionary file, comments and the FoamFile header block skipped
std::map<std::string, std::string> facadeReadDict(const std::string &file)
{
    std::map<std::string, std::string> out;
    std::ifstream in(file);
    std::stringstream ss;
    ss << in.rdbuf();
    std::string s = ss.str(), clean;
    for (size_t i = 0; i < s.size();)
    {
        if (s.compare(i, 2, "//") == 0)
        {
            while (i < s.size() && s[i] != '\n')
                ++i;
        }
        else if (s.compare(i, 2, "/*") == 0)
        {
            const size_t e = s.find("*/", i + 2);
            i = (e == std::string::npos) ? s.size() : e + 2;
        }
        else
            clean += s[i++];
    }
    int depth = 0;
    std::string stmt;
    for (char ch : clean)
    {
        if (ch == '{')
            ++depth;
        else if (ch == '}')
        {
            --depth;
            stmt.clear();
        }
        else if (depth == 0)
        {
            if (ch == ';')
            {
                std::istringstream is(stmt);
                std::string k, v;
                is >> k >> v;
                if (!k.empty())
                    out[k] = v;
                stmt.clear();
            }
            else
                stmt += ch;
        }
    }
    return out;
}
// numeric directory names of a case, ascending
std::vector<std::pair<double, std::string>> facadeFindTimes(const std::string &caseDir)
{
    std::vector<std::pair<double, std::string>> out;
    DIR *d = opendir(caseDir.c_str());
    if (!d)
        return out;
    while (dirent *e = readdir(d))
    {
        const std::string n = e->d_name;
        char *end = nullptr;
        const double v = strtod(n.c_str(), &end);
        if (!n.empty() && end && *end == 0 && (isdigit((unsigned char)n[0]) || n[0] == '-' || n[0] == '.'))
            out.push_back({v, n});
    }
    closedir(d);
    std::sort(out.begin(), out.end());
    return out;
}
} // namespace Foam

using namespace Foam; // what OpenFOAM application sources assume

#include "/root/reference/src/smoothMesh.C"
