// parmparse.hpp -- the slice of amrex::ParmParse the two tools use (AMReX.cpp:404-445, AMReX_ParmParse.H):
// arguments are "key=value ..." on the command line, optionally preceded by an inputs file (first argument without
// '='); a value may have several tokens; '#' starts a comment in files; later definitions override earlier ones.
#pragma once
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

class ParmParse {
public:
    ParmParse(int argc, char** argv) {
        int first = 1;
        if (argc > 1 && std::string(argv[1]).find('=') == std::string::npos) {
            std::ifstream f(argv[1]);
            if (!f) throw std::runtime_error(std::string("ParmParse: cannot open inputs file ") + argv[1]);
            std::stringstream ss;
            std::string line;
            while (std::getline(f, line)) {
                auto h = line.find('#');
                if (h != std::string::npos) line.erase(h);
                ss << line << '\n';
            }
            parse(ss.str());
            first = 2;
        }
        std::string cmd;
        for (int i = first; i < argc; ++i) { cmd += argv[i]; cmd += ' '; }
        parse(cmd);
    }
    bool contains(const std::string& k) const { return tab_.count(k) > 0; }
    int countval(const std::string& k) const { auto it = tab_.find(k); return it == tab_.end() ? 0 : (int)it->second.size(); }
    template <class T> bool query(const std::string& k, T& v, int idx = 0) const {
        auto it = tab_.find(k);
        if (it == tab_.end() || idx >= (int)it->second.size()) return false;
        conv(it->second[idx], v);
        return true;
    }
    template <class T> void get(const std::string& k, T& v, int idx = 0) const {
        if (!query(k, v, idx)) { std::cerr << "ParmParse::get: " << k << " not found\n"; std::exit(1); }   // amrex::Abort
    }
    template <class T> bool queryarr(const std::string& k, std::vector<T>& v, int start, int n) const {
        auto it = tab_.find(k);
        if (it == tab_.end()) return false;
        for (int i = 0; i < n && start + i < (int)it->second.size(); ++i) {
            if ((int)v.size() <= i) v.resize(i + 1);
            conv(it->second[start + i], v[i]);
        }
        return true;
    }
private:
    std::map<std::string, std::vector<std::string>> tab_;
    static void conv(const std::string& s, std::string& v) { v = s; }
    static void conv(const std::string& s, int& v) { v = std::atoi(s.c_str()); }
    static void conv(const std::string& s, double& v) { v = std::atof(s.c_str()); }
    static void conv(const std::string& s, bool& v) { v = (s == "1" || s == "true" || s == "T" || s == "t" || s == "True"); }
    void parse(const std::string& text) {
        // tokenise, keeping "=" as its own token and honouring double quotes
        std::vector<std::string> tok;
        std::string cur;
        bool q = false;
        auto flush = [&] { if (!cur.empty()) { tok.push_back(cur); cur.clear(); } };
        for (char c : text) {
            if (c == '"') { q = !q; continue; }
            if (!q && (c == ' ' || c == '\t' || c == '\n' || c == '\r')) { flush(); continue; }
            if (!q && c == '=') { flush(); tok.push_back("="); continue; }
            cur += c;
        }
        flush();
        size_t i = 0;
        while (i < tok.size()) {
            if (i + 1 < tok.size() && tok[i + 1] == "=") {
                std::string key = tok[i];
                i += 2;
                std::vector<std::string> vals;
                while (i < tok.size() && !(i + 1 < tok.size() && tok[i + 1] == "=")) vals.push_back(tok[i++]);
                tab_[key] = vals;
            } else {
                ++i;
            }
        }
    }
};
