// Known-answer tests for statistics.hpp: the cases of the reference's test/test_statistics.cpp
// (TestVariable :168-260, TestCounter :288-310, TestPeak :336-416, TestStatisticsRegistry :466-608),
// restated on this implementation.  Prints "statistics_test: ok" and exits 0, or names the failed check.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "statistics.hpp"

namespace st = mlsb::statistics;

#define CHECK(cond) \
    do { if (!(cond)) { std::fprintf(stderr, "statistics_test: %s:%d: %s\n", __FILE__, __LINE__, #cond); std::exit(1); } } while (0)

template <typename E, typename F>
static bool throws(F f)
{
    try { f(); }
    catch (const E &) { return true; }
    catch (...) { return false; }
    return false;
}

template <typename T>
static std::string str(const T &v)
{
    std::ostringstream o;
    o << v;
    return o.str();
}

static void testVariable()
{
    st::Variable stat0("stat0"), stat1("stat1"), stat2("stat2"), stat2s("stat2s");
    stat1.add(1.0);
    stat2.add(2.0);
    stat2.add(3.0);
    stat2s.add(4.5);
    stat2s.add(4.5);
    CHECK(stat0.getName() == "stat0");
    CHECK(stat1.getSum() == 1.0 && stat2.getSum() == 5.0 && stat2s.getSum() == 9.0);
    CHECK(throws<std::length_error>([&] { stat0.getMean(); }));
    CHECK(stat1.getMean() == 1.0 && stat2.getMean() == 2.5 && stat2s.getMean() == 4.5);
    CHECK(throws<std::length_error>([&] { stat0.getVariance(); }));
    CHECK(throws<std::length_error>([&] { stat1.getVariance(); }));
    CHECK(stat2.getVariance() == 0.5 && stat2s.getVariance() == 0.0);
    CHECK(throws<std::length_error>([&] { stat0.getStddev(); }));
    CHECK(throws<std::length_error>([&] { stat1.getStddev(); }));
    CHECK(std::fabs(stat2.getStddev() - 0.707106781186548) < 1e-12 && stat2s.getStddev() == 0.0);
    CHECK(stat0.getNumSamples() == 0 && stat1.getNumSamples() == 1 && stat2.getNumSamples() == 2 && stat2s.getNumSamples() == 2);
    CHECK(str(stat0) == "stat0: [0]");
    CHECK(str(stat1) == "stat1: 1 : 1 [1]");
    CHECK(str(stat2) == "stat2: 5 : 2.5 +/- 0.707107 [2]");
    CHECK(str(stat2s) == "stat2s: 9 : 4.5 +/- 0 [2]");
    std::unique_ptr<st::Statistic> copy = stat2.clone();
    CHECK(str(*copy) == str(stat2) && dynamic_cast<st::Variable *>(copy.get()) != nullptr);
    stat1.merge(stat2);
    CHECK(stat1.getNumSamples() == 3 && stat1.getSum() == 6.0 && stat1.getVariance() == 1.0);
    st::Counter other("c");
    CHECK(throws<std::bad_cast>([&] { stat1.merge(other); }));
}

static void testCounter()
{
    st::Counter counter("counter");
    counter.add(100);
    CHECK(counter.getTotal() == 100);
    CHECK(str(counter) == "counter: 100");
    counter.add(50);
    CHECK(counter.getTotal() == 150);
    counter.add();
    CHECK(counter.getTotal() == 151);
    st::Counter more("counter");
    more.add(9);
    counter.merge(more);
    CHECK(counter.getTotal() == 160);
    CHECK(str(*counter.clone()) == "counter: 160");
}

static void testPeak()
{
    st::Peak peak("peak"), empty("empty");
    CHECK(empty.get() == 0 && empty.getMax() == 0 && str(empty) == "empty: 0");
    peak = -100;
    CHECK(peak.get() == -100 && peak.getMax() == 0);
    peak = 1234567890;
    CHECK(peak.get() == 1234567890 && peak.getMax() == 1234567890);
    peak = 123456;
    CHECK(peak.get() == 123456 && peak.getMax() == 1234567890);

    st::Peak p2("peak");
    p2 = 100;
    p2 += 50;
    CHECK(p2.get() == 150 && p2.getMax() == 150);
    p2 += -200;
    CHECK(p2.get() == -50 && p2.getMax() == 150);
    p2 -= -200;
    CHECK(p2.get() == 150 && p2.getMax() == 150);
    p2 -= 200;
    CHECK(p2.get() == -50 && p2.getMax() == 150);
    p2 = 123;
    CHECK(str(p2) == "peak: 150");
    st::Peak p3("peak");
    p3 = 123;
    CHECK(str(p3) == "peak: 123");
    p3.merge(p2);
    CHECK(p3.getMax() == 150 && p3.get() == 123);
    p2.merge(p3);
    CHECK(p2.getMax() == 150);
}

static void fill(st::Registry &r)
{
    r.getStatistic<st::Variable>("stat0");
    r.getStatistic<st::Variable>("stat1").add(1.0);
    st::Variable &stat3 = r.getStatistic<st::Variable>("stat3");
    stat3.add(2.0);
    stat3.add(4.0);
    stat3.add(6.0);
    r.getStatistic<st::Counter>("counter").add(100);
}

static void testRegistry()
{
    CHECK(&st::Registry::getInstance() == &st::Registry::getInstance());
    st::Registry registry;
    fill(registry);
    CHECK(registry.getStatistic<st::Variable>("stat1").getNumSamples() == 1);
    CHECK(registry.getStatistic<st::Variable>("new").getNumSamples() == 0);
    CHECK(throws<std::bad_cast>([&] { registry.getStatistic<st::Variable>("counter"); }));
    CHECK(&registry.getStatistic<st::Counter>("counter") == &registry.getStatistic<st::Counter>("counter"));

    st::Registry r2;
    fill(r2);
    CHECK(str(r2) == "counter: 100\nstat0: [0]\nstat1: 1 : 1 [1]\nstat3: 12 : 4 +/- 2 [3]\n");

    // iteration in name order, mutable
    std::vector<std::string> names;
    r2.forEach([&](st::Statistic &s) {
        names.push_back(s.getName());
        if (st::Variable *v = dynamic_cast<st::Variable *>(&s)) v->add(1.0);
        else if (st::Counter *c = dynamic_cast<st::Counter *>(&s)) c->add(1);
    });
    CHECK(names.size() == 4 && names[0] == "counter" && names[1] == "stat0" && names[2] == "stat1" && names[3] == "stat3");
    CHECK(r2.getStatistic<st::Variable>("stat0").getNumSamples() == 1 && r2.getStatistic<st::Variable>("stat1").getNumSamples() == 2);
    CHECK(r2.getStatistic<st::Variable>("stat3").getNumSamples() == 4 && r2.getStatistic<st::Counter>("counter").getTotal() == 101);

    // merge: same names are summed, new names are copied
    st::Registry mine, other;
    fill(mine);
    other.getStatistic<st::Variable>("stat0").add(12.0);
    other.getStatistic<st::Counter>("counter").add(17);
    other.getStatistic<st::Variable>("extra").add(3.0);
    mine.merge(other);
    CHECK(str(mine) == "counter: 117\nextra: 3 : 3 [1]\nstat0: 12 : 12 [1]\nstat1: 1 : 1 [1]\nstat3: 12 : 4 +/- 2 [3]\n");
    other.getStatistic<st::Variable>("extra").add(1.0);   // the copy is independent of its source
    CHECK(mine.getStatistic<st::Variable>("extra").getNumSamples() == 1);
}

static void testThreadsAndTimer()
{
    st::Registry r;
    std::vector<std::thread> threads;
    for (int t = 0; t < 8; t++)
        threads.emplace_back([&r] {
            for (int i = 0; i < 10000; i++)
            {
                r.getStatistic<st::Counter>("hits").add();
                r.getStatistic<st::Variable>("v").add(2.0);
                st::Peak &p = r.getStatistic<st::Peak>("live");
                p += 3;
                p -= 3;
            }
        });
    for (std::thread &t : threads) t.join();
    CHECK(r.getStatistic<st::Counter>("hits").getTotal() == 80000);
    CHECK(r.getStatistic<st::Variable>("v").getNumSamples() == 80000 && r.getStatistic<st::Variable>("v").getMean() == 2.0);
    CHECK(r.getStatistic<st::Peak>("live").get() == 0 && r.getStatistic<st::Peak>("live").getMax() >= 3 && r.getStatistic<st::Peak>("live").getMax() <= 24);
    st::Variable &timed = r.getStatistic<st::Variable>("timed");
    {
        st::Timer timer(timed);
        CHECK(timer.getElapsed() >= 0.0);
    }
    CHECK(timed.getNumSamples() == 1 && timed.getSum() >= 0.0);
}

int main()
{
    testVariable();
    testCounter();
    testPeak();
    testRegistry();
    testThreadsAndTimer();
    std::puts("statistics_test: ok");
    return 0;
}
