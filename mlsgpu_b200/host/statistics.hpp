// Statistics registry of the command-line driver: named counters, sampled variables and
// high-water marks, collected from any thread and written as "name: value" lines in name order.
// Mirrors the interface and the output format of the reference's Statistics namespace
// (src/statistics.h:132-262 and :316-420; write formats src/statistics.cpp:99-101, :175-182,
// :206-208; registry stream :340-349), so a --statistics file reads the same:
//
//   counter      "files.splats: 1000000"
//   variable     "run.time: 7.97 : 7.97 [1]"   sum : mean [+/- sample stddev] [samples]
//   peak         "mem.splats: 3200000000"      the highest value the quantity reached
//
// No serialization (the reference needs it for its MPI gather; merge() is what a gather would use).
#ifndef MLSB200_STATISTICS_HPP
#define MLSB200_STATISTICS_HPP

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <ostream>
#include <stdexcept>
#include <string>
#include <typeinfo>

namespace mlsb
{
namespace statistics
{

class Statistic
{
public:
    explicit Statistic(const std::string &name) : name(name) {}
    virtual ~Statistic() {}
    Statistic(const Statistic &) = delete;
    Statistic &operator=(const Statistic &) = delete;

    const std::string &getName() const { return name; }

    /// Adds the samples of `other`; std::bad_cast if it is a different kind of statistic.
    virtual void merge(const Statistic &other) = 0;
    /// A new statistic of the same kind, name and contents.
    virtual std::unique_ptr<Statistic> clone() const = 0;

    friend std::ostream &operator<<(std::ostream &o, const Statistic &s)
    {
        std::lock_guard<std::mutex> lock(s.mutex);
        o << s.name << ": ";
        s.write(o);
        return o;
    }

protected:
    mutable std::mutex mutex;
    virtual void write(std::ostream &o) const = 0;   // value only; caller holds the lock

private:
    const std::string name;
};

/// A running total (src/statistics.h:132-157).
class Counter : public Statistic
{
public:
    explicit Counter(const std::string &name) : Statistic(name) {}
    void add(unsigned long long incr = 1ULL)
    {
        std::lock_guard<std::mutex> lock(mutex);
        total += incr;
    }
    unsigned long long getTotal() const
    {
        std::lock_guard<std::mutex> lock(mutex);
        return total;
    }
    void merge(const Statistic &other) override
    {
        const unsigned long long t = dynamic_cast<const Counter &>(other).getTotal();
        add(t);
    }
    std::unique_ptr<Statistic> clone() const override
    {
        std::unique_ptr<Counter> c(new Counter(getName()));
        c->total = getTotal();
        return std::unique_ptr<Statistic>(c.release());
    }

protected:
    void write(std::ostream &o) const override { o << total; }

private:
    unsigned long long total = 0;
};

/// Samples of a quantity: count, mean, sample standard deviation (src/statistics.h:164-209).
class Variable : public Statistic
{
public:
    explicit Variable(const std::string &name) : Statistic(name) {}
    void add(double value)
    {
        std::lock_guard<std::mutex> lock(mutex);
        sum += value;
        sum2 += value * value;
        n++;
    }
    /// Adds `count` samples given by their sum and sum of squares (per-kernel times gathered on the device side).
    void addSamples(double sampleSum, double sampleSum2, unsigned long long count)
    {
        std::lock_guard<std::mutex> lock(mutex);
        sum += sampleSum;
        sum2 += sampleSum2;
        n += count;
    }
    unsigned long long getNumSamples() const
    {
        std::lock_guard<std::mutex> lock(mutex);
        return n;
    }
    double getSum() const
    {
        std::lock_guard<std::mutex> lock(mutex);
        return sum;
    }
    double getMean() const
    {
        std::lock_guard<std::mutex> lock(mutex);
        if (n < 1) throw std::length_error("Cannot compute mean without at least 1 sample");
        return sum / (double) n;
    }
    double getVariance() const
    {
        std::lock_guard<std::mutex> lock(mutex);
        return variance();
    }
    double getStddev() const { return std::sqrt(getVariance()); }
    void merge(const Statistic &other) override
    {
        const Variable &v = dynamic_cast<const Variable &>(other);
        double s, s2;
        unsigned long long c;
        {
            std::lock_guard<std::mutex> lock(v.mutex);
            s = v.sum, s2 = v.sum2, c = v.n;
        }
        addSamples(s, s2, c);
    }
    std::unique_ptr<Statistic> clone() const override
    {
        std::unique_ptr<Variable> v(new Variable(getName()));
        v->merge(*this);
        return std::unique_ptr<Statistic>(v.release());
    }

protected:
    void write(std::ostream &o) const override
    {
        if (n >= 1) o << sum << " : " << sum / (double) n << ' ';
        if (n >= 2) o << "+/- " << std::sqrt(variance()) << ' ';
        o << "[" << n << "]";
    }

private:
    double sum = 0.0, sum2 = 0.0;
    unsigned long long n = 0;
    double variance() const
    {
        if (n < 2) throw std::length_error("Cannot compute variance without at least 2 samples");
        // rounding may push the difference below zero
        return std::max((sum2 - sum * sum / (double) n) / (double) (n - 1), 0.0);
    }
};

/// A quantity that goes up and down, of which the maximum is reported (src/statistics.h:216-262).
class Peak : public Statistic
{
public:
    typedef std::int64_t value_type;
    explicit Peak(const std::string &name) : Statistic(name) {}
    Peak &operator+=(value_type x)
    {
        std::lock_guard<std::mutex> lock(mutex);
        set(current + x);
        return *this;
    }
    Peak &operator-=(value_type x)
    {
        std::lock_guard<std::mutex> lock(mutex);
        set(current - x);
        return *this;
    }
    Peak &operator=(value_type x)
    {
        std::lock_guard<std::mutex> lock(mutex);
        set(x);
        return *this;
    }
    value_type get() const
    {
        std::lock_guard<std::mutex> lock(mutex);
        return current;
    }
    value_type getMax() const
    {
        std::lock_guard<std::mutex> lock(mutex);
        return peak;
    }
    /// The larger of the two maxima; the current value is this one's (src/statistics.cpp:260-265).
    void merge(const Statistic &other) override
    {
        const value_type m = dynamic_cast<const Peak &>(other).getMax();
        std::lock_guard<std::mutex> lock(mutex);
        peak = std::max(peak, m);
    }
    std::unique_ptr<Statistic> clone() const override
    {
        std::unique_ptr<Peak> p(new Peak(getName()));
        {
            std::lock_guard<std::mutex> lock(mutex);
            p->current = current;
            p->peak = peak;
        }
        return std::unique_ptr<Statistic>(p.release());
    }

protected:
    void write(std::ostream &o) const override { o << peak; }

private:
    value_type current = 0, peak = 0;
    void set(value_type x)
    {
        current = x;
        peak = std::max(peak, current);
    }
};

/// Statistics by name (src/statistics.h:316-420).  getStatistic<T>() creates on first use.
class Registry
{
public:
    typedef std::map<std::string, std::unique_ptr<Statistic>> map_type;

    /// The process-wide registry.
    static Registry &getInstance()
    {
        static Registry instance;
        return instance;
    }

    template <typename T>
    T &getStatistic(const std::string &name)
    {
        std::lock_guard<std::mutex> lock(mutex);
        map_type::iterator it = stats.find(name);
        if (it == stats.end()) it = stats.emplace(name, std::unique_ptr<Statistic>(new T(name))).first;
        return dynamic_cast<T &>(*it->second);   // std::bad_cast if the name is taken by another kind
    }

    bool has(const std::string &name) const
    {
        std::lock_guard<std::mutex> lock(mutex);
        return stats.count(name) != 0;
    }

    std::size_t size() const
    {
        std::lock_guard<std::mutex> lock(mutex);
        return stats.size();
    }

    /// Calls f(statistic) in name order.
    template <typename F>
    void forEach(F f)
    {
        std::lock_guard<std::mutex> lock(mutex);
        for (map_type::value_type &e : stats) f(*e.second);
    }
    template <typename F>
    void forEach(F f) const
    {
        std::lock_guard<std::mutex> lock(mutex);
        for (const map_type::value_type &e : stats) f(static_cast<const Statistic &>(*e.second));
    }

    /// Statistics of `other` are merged into those of the same name, or cloned if new here.
    void merge(const Registry &other)
    {
        if (&other == this) return;
        std::lock_guard<std::mutex> theirs(other.mutex);
        std::lock_guard<std::mutex> lock(mutex);
        for (const map_type::value_type &e : other.stats)
        {
            map_type::iterator it = stats.find(e.first);
            if (it == stats.end())
                stats.emplace(e.first, e.second->clone());
            else
                it->second->merge(*e.second);
        }
    }

    friend std::ostream &operator<<(std::ostream &o, const Registry &r)
    {
        std::lock_guard<std::mutex> lock(r.mutex);
        for (const map_type::value_type &e : r.stats) o << *e.second << '\n';
        return o;
    }

private:
    mutable std::mutex mutex;
    map_type stats;
};

template <typename T>
inline T &getStatistic(const std::string &name)
{
    return Registry::getInstance().getStatistic<T>(name);
}

/// Adds the time between construction and destruction to a variable (Statistics::Timer).
class Timer
{
public:
    explicit Timer(const std::string &name) : stat(getStatistic<Variable>(name)), start(clock::now()) {}
    explicit Timer(Variable &stat) : stat(stat), start(clock::now()) {}
    Timer(const Timer &) = delete;
    Timer &operator=(const Timer &) = delete;
    double getElapsed() const { return std::chrono::duration<double>(clock::now() - start).count(); }
    ~Timer() { stat.add(getElapsed()); }

private:
    typedef std::chrono::steady_clock clock;
    Variable &stat;
    const clock::time_point start;
};

}  // namespace statistics
}  // namespace mlsb

#endif
